"""Synthetic weight provider (replaces the reference's Hugging Face loader, hf_model_utils.py:135-287).

The build box has no network, so BASELINE.json's configs run on synthetic bf16-exact tensors
of DeepSeek-R1 weight shapes.  Values follow the survey's recipe:
``torch.randn(shape, generator=manual_seed(s)) * 0.02 -> bfloat16`` on the CPU generator, which
is deterministic for a given torch build, so fixtures made in the build container can be
regenerated bit-identically on the GPU box (same image).
"""
from __future__ import annotations

import numpy as np
import torch

# name -> [rows, cols]; layer-0 attention, dense MLP and one routed expert.
DEEPSEEK_R1_SHAPES: dict[str, tuple[int, int]] = {
    "model.layers.0.self_attn.q_a_proj.weight": (1536, 7168),
    "model.layers.0.self_attn.q_b_proj.weight": (24576, 1536),
    "model.layers.0.self_attn.kv_a_proj_with_mqa.weight": (576, 7168),
    "model.layers.0.self_attn.kv_b_proj.weight": (32768, 512),
    "model.layers.0.self_attn.o_proj.weight": (7168, 16384),
    "model.layers.0.mlp.gate_proj.weight": (18432, 7168),
    "model.layers.0.mlp.up_proj.weight": (18432, 7168),
    "model.layers.0.mlp.down_proj.weight": (7168, 18432),
}
ATTN_NAMES = [n for n in DEEPSEEK_R1_SHAPES if ".self_attn." in n]
MLP_NAMES = [n for n in DEEPSEEK_R1_SHAPES if ".mlp." in n]
EXPERT_SHAPES = {"gate_proj": (2048, 7168), "up_proj": (2048, 7168), "down_proj": (7168, 2048)}


def expert_tensor_list(num_experts: int = 256, layer: int = 3) -> list[tuple[str, tuple[int, int]]]:
    """Config 5: one MoE layer = num_experts x (gate, up, down)."""
    out = []
    for e in range(num_experts):
        for proj, shp in EXPERT_SHAPES.items():
            out.append((f"model.layers.{layer}.mlp.experts.{e}.{proj}.weight", shp))
    return out


def randn_bf16_cpu(shape, seed: int, scale: float = 0.02) -> torch.Tensor:
    """CPU bf16 tensor; the survey's `randn * 0.02 -> bf16` recipe."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return (torch.randn(tuple(shape), generator=g, dtype=torch.float32) * scale).to(torch.bfloat16)


def fp8_checkpoint_cpu(shape, seed: int, block=(128, 128), scale: float = 0.02):
    """A 2-D weight as an fp8 checkpoint stores it (what hf_model_utils.py:199-215 dequantizes): randn * scale quantized per
    `block` to e4m3fn with inverse scale = block amax / 448.  -> (uint8 [rows, cols] e4m3fn bytes, float32 [ceil(rows / block
    rows), ceil(cols / block cols)] inverse scales)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    rows, cols = (int(v) for v in shape)
    x = torch.randn((rows, cols), generator=g, dtype=torch.float32) * scale
    br, bc = block
    sr, scn = -(-rows // br), -(-cols // bc)
    pad = torch.zeros((sr * br, scn * bc), dtype=torch.float32)
    pad[:rows, :cols] = x
    amax = pad.reshape(sr, br, scn, bc).abs().amax(dim=(1, 3)).clamp_min(1e-12)
    inv = (amax / 448.0).to(torch.float32)
    q = (pad.reshape(sr, br, scn, bc) / inv[:, None, :, None]).reshape(sr * br, scn * bc)[:rows, :cols]
    return q.to(torch.float8_e4m3fn).view(torch.uint8).contiguous(), inv.contiguous()


def randn_f32_np(shape, seed: int, scale: float = 0.02) -> np.ndarray:
    """Same values as float32 NumPy (bf16-exact)."""
    return randn_bf16_cpu(shape, seed, scale).to(torch.float32).numpy()


def heterogeneous_f32_np(shape, seed: int) -> np.ndarray:
    """Heavy-tailed family: per-32x32-tile log-normal scale (sigma 0.7) and 0.1 % x20 outliers,
    then rounded to bf16.  Exercises wide exponent spreads and mixed greedy decisions."""
    rng = np.random.default_rng(seed)
    shape = tuple(int(s) for s in shape)
    x = rng.standard_normal(shape).astype(np.float32) * np.float32(0.02)
    if x.ndim >= 2:
        r, c = int(np.prod(shape[:-1])), shape[-1]
        s = np.exp(rng.standard_normal((-(-r // 32), -(-c // 32))) * 0.7).astype(np.float32)
        s = np.repeat(np.repeat(s, 32, axis=0), 32, axis=1)[:r, :c]
        x = (x.reshape(r, c) * s).reshape(shape)
    out = rng.random(shape) < 1e-3
    x = np.where(out, x * np.float32(20.0), x).astype(np.float32)
    return torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()


def device_randn_bf16(shape, seed: int, device, scale: float = 0.02) -> torch.Tensor:
    """On-device generation for the large configs (values differ from the CPU generator;
    used only where the oracle checks size-independent properties or a D2H subset)."""
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return (torch.randn(tuple(shape), generator=g, dtype=torch.float32, device=device) * scale).to(torch.bfloat16)
