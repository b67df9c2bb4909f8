"""Where `wq` and the sweep get their tensors (replaces the network half of hf_model_utils.py:135-287).

The reference resolves a Hugging Face repo, downloads safetensors shards, dequantizes fp8 blocks and keeps every tensor as
float32 ``.npy`` under ``<cache-dir>/tensor-fp32/<repo--rev--sha1[:12]>/<tensor--sha1[:12]>.npy`` (hf_model_utils.py:114-132,
245-287).  This build has no network: a repo is served from that same on-disk cache when it is populated (so a cache
filled by the reference - or by a test - is a drop-in source), and the pseudo-repo ``synthetic`` (or any repo whose cache
is empty while ``allow_synthetic`` is set) maps to the synthetic DeepSeek-R1 shapes of ``synthetic.py``.
"""
from __future__ import annotations

import hashlib
import re
from dataclasses import dataclass, field
from pathlib import Path
from urllib.parse import urlparse

import numpy as np

from . import synthetic

SYNTHETIC_REPO = "synthetic/DeepSeek-R1-shapes"


def normalize_repo_id(raw: str) -> str:
    """`org/name`, or a huggingface.co URL of a model repo (hf_model_utils.py:25-57)."""
    value = raw.strip()
    if not value:
        raise ValueError("Empty repo value.")
    if "://" not in value:
        return value.strip("/")
    u = urlparse(value)
    host = u.netloc.lower()
    host = host[4:] if host.startswith("www.") else host
    if host not in ("huggingface.co", "hf.co"):
        raise ValueError(f"Unsupported host: {u.netloc}")
    parts = [p for p in u.path.split("/") if p]
    if not parts:
        raise ValueError("URL path does not contain a repo id.")
    if parts[0] in ("models", "model"):
        parts = parts[1:]
    elif parts[0] in ("datasets", "spaces"):
        raise ValueError("Only model repos are supported.")
    for i, p in enumerate(parts):
        if p in ("tree", "blob", "resolve", "commit", "discussions"):
            parts = parts[:i]
            break
    return "/".join(parts[:2]) if len(parts) >= 2 else parts[0]


def safe_repo_revision_key(repo_id: str, revision: str) -> str:
    """hf_model_utils.py:114-118."""
    digest = hashlib.sha1(f"{repo_id}@{revision}".encode("utf-8")).hexdigest()[:12]
    return f"{repo_id.replace('/', '__')}--{re.sub(r'[^A-Za-z0-9._-]+', '_', revision)}--{digest}"


def safe_tensor_key(tensor_name: str) -> str:
    """hf_model_utils.py:121-126."""
    digest = hashlib.sha1(tensor_name.encode("utf-8")).hexdigest()[:12]
    safe = re.sub(r"[^A-Za-z0-9._-]+", "_", tensor_name).strip("_") or "tensor"
    return f"{safe}--{digest}"


def filter_tensor_names(names, query):
    """Substring, or dotted prefix path (hf_model_utils.py:60-77)."""
    if not query or not query.strip():
        return sorted(names)
    q = query.strip()
    if "." in q:
        qp = [p.lower() for p in q.split(".") if p]
        return sorted(n for n in names if n.lower().split(".")[: len(qp)] == qp)
    return sorted(n for n in names if q.lower() in n.lower())


def resolve_format_list(values, supported):
    """hf_model_utils.py:317-335."""
    if not values:
        return list(supported)
    out = []
    for raw in values:
        v = raw.strip().lower()
        if v == "all":
            out += [s for s in supported if s not in out]
            continue
        if v not in supported:
            raise ValueError(f"Unsupported format '{raw}'. Supported: {', '.join(supported)}, all")
        if v not in out:
            out.append(v)
    return out


@dataclass
class TensorIndex:
    """The part of hf_model_utils.ModelIndex (:103-111) the analysis loop uses."""
    repo_id: str
    revision: str
    cache_dir: Path
    tensor_to_file: dict = field(default_factory=dict)      # name -> .npy path, or "synthetic"
    synthetic_seed: int = 1000

    def fp32_cache_dir(self) -> Path:
        return Path(self.cache_dir) / "tensor-fp32" / safe_repo_revision_key(self.repo_id, self.revision)

    def cache_file(self, name: str) -> Path:
        return self.fp32_cache_dir() / f"{safe_tensor_key(name)}.npy"

    def load_fp32(self, name: str) -> np.ndarray:
        src = self.tensor_to_file[name]
        if src == "synthetic":
            i = list(self.tensor_to_file).index(name)
            return synthetic.randn_f32_np(synthetic.DEEPSEEK_R1_SHAPES[name], self.synthetic_seed + i)
        return np.load(src)


def build_tensor_index(repo_or_url: str, revision: str = "main", cache_dir="data/hf-cache", synthetic_seed: int = 1000) -> TensorIndex:
    repo_id = normalize_repo_id(repo_or_url)
    idx = TensorIndex(repo_id=repo_id, revision=revision, cache_dir=Path(cache_dir), synthetic_seed=synthetic_seed)
    d = idx.fp32_cache_dir()
    names = {}
    if d.is_dir():
        for f in sorted(d.glob("*.npy")):
            stem = f.name[:-4]
            if "--" not in stem:
                continue
            name, digest = stem.rsplit("--", 1)
            # the cache key keeps tensor names made of [A-Za-z0-9._-] verbatim; the digest tells whether it did
            if hashlib.sha1(name.encode("utf-8")).hexdigest()[:12] == digest:
                names[name] = str(f)
    if names:
        idx.tensor_to_file = names
        return idx
    if repo_id.split("/")[0].lower() == "synthetic" or repo_id.lower() == "synthetic":
        idx.tensor_to_file = {n: "synthetic" for n in synthetic.DEEPSEEK_R1_SHAPES}
        return idx
    raise RuntimeError(
        f"No cached float32 tensors for {repo_id}@{revision} under {d} and no network access in this build: "
        f"populate that directory (the reference's own cache layout) or use the repo name 'synthetic'.")


def resolve_selected_tensors(index: TensorIndex, filter_query) -> list[str]:
    """hf_model_utils.py:290-301."""
    names = list(index.tensor_to_file)
    weight_like = [n for n in names if "weight" in n.lower() and not n.lower().endswith("_scale_inv")]
    sel = filter_tensor_names(weight_like if weight_like else names, filter_query)
    if not sel:
        sel = filter_tensor_names(names, filter_query)
    if not sel:
        raise RuntimeError("No tensors matched the filter query.")
    return sel
