"""Shared helpers for the test-suite (CPU and GPU)."""
import os

import torch

from magvit2_pytorch_b200 import VideoTokenizer
from oracle import weights as W
from oracle.restated import OracleTokenizer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

README_LAYERS = (
    "residual", "compress_space", ("consecutive_residual", 2), "compress_space",
    ("consecutive_residual", 2), "linear_attend_space", "compress_space",
    ("consecutive_residual", 2), "attend_space", "compress_time",
    ("consecutive_residual", 2), "compress_time", ("consecutive_residual", 2), "attend_time",
)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, f"{name}.pt"), map_location="cpu", weights_only=False)


def build_product(kwargs, wseed=0):
    """Product model (CPU, fp32) with the deterministic synthetic weights."""
    torch.manual_seed(0)
    m = VideoTokenizer(**kwargs)
    W.fill_state_dict_(m, wseed)
    m.eval()
    return m


def build_oracle(model, kwargs, dtype=torch.float32):
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    return OracleTokenizer(sd, dtype=dtype, **kwargs)


def build_oracle_from_golden(g, dtype=torch.float32):
    """Oracle for a golden whose spec the product does not construct (conditioned layers): the reference's state_dict
    layout is stored in the golden, the synthetic weights are rebuilt from it."""
    sd = {k: W.synth_tensor(k, shp, g["wseed"]) for k, shp in g["sd_shapes"].items()}
    sd.update({k: v.clone() for k, v in g["sd_buffers"].items()})
    return OracleTokenizer(sd, dtype=dtype, **g["kwargs"])


def golden_video(g):
    b, c, t, s = g["video_shape"][:4]
    return W.synth_video(b, c, t, s, seed=g["vseed"])


def sample_like_golden(t, g):
    cs, ss = g.get("tap_strides", (7, 5))
    return t[:, ::cs, :, ::ss, ::ss]
