# Tile-stat kernel variants of round 2 (libraries prebuilt as libqa_<name>.so from -D flags in qa_stats.cu), o_proj-size tensor.
P=quantization_analysis_b200
cp $P/libqa_b200.so /tmp/keep.so
python profiles/stats_time.py 2 base
for v in "$@"; do
  cp $P/libqa_$v.so $P/libqa_b200.so
  python profiles/stats_time.py 2 $v
done
cp /tmp/keep.so $P/libqa_b200.so
