"""Pins oracle/restated.py (the CPU restatement used as the checker on the GPU box) against
golden vectors produced by the unmodified reference (oracle/make_golden.py), and checks the
behavioural properties SURVEY.md section 4 lists.  CPU only."""
import pytest
import torch

from tests.util import build_oracle, build_product, golden_video, load_golden, sample_like_golden

SMALL = ["cfg1", "mini", "mini_fsq"]


@pytest.mark.parametrize("name", SMALL)
def test_restated_matches_reference_golden(name):
    g = load_golden(name)
    model = build_product(g["kwargs"], g["wseed"])
    orc = build_oracle(model, g["kwargs"])
    video = golden_video(g)
    taps = {}
    codes = orc.tokenize(video, taps=taps)
    assert codes.dtype == g["codes"].dtype
    assert torch.equal(codes, g["codes"]), "restated oracle codes differ from the reference's"
    for k, ref in g["taps"].items():
        if k.startswith("enc") or k == "conv_in":
            got = sample_like_golden(taps[k], g)
            assert torch.allclose(got, ref, atol=2e-5, rtol=1e-5), k
    dtaps = {}
    recon = orc.decode_from_code_indices(codes, taps=dtaps)
    assert torch.allclose(recon, g["recon"], atol=2e-5, rtol=1e-5)
    for k, ref in g["taps"].items():
        if k.startswith("dec"):
            assert torch.allclose(sample_like_golden(dtaps[k], g), ref, atol=2e-5, rtol=1e-5), k


def test_restated_matches_reference_golden_readme():
    """BASELINE configs[1] (README config), one clip: ~6 s of CPU."""
    g = load_golden("readme")
    model = build_product(g["kwargs"], g["wseed"])
    orc = build_oracle(model, g["kwargs"])
    video = golden_video(g)
    codes, pre = orc.tokenize(video, return_presign=True)
    assert torch.equal(codes, g["codes"])
    assert torch.allclose(pre, g["presign"], atol=1e-4)
    recon = orc.decode_from_code_indices(codes)
    assert torch.allclose(recon[:, :, :, ::4, ::4], g["recon_sample"], atol=5e-5, rtol=1e-5)
    assert torch.allclose(recon.mean(dim=(3, 4)), g["recon_mean"], atol=1e-5)


def test_readme_roundtrip_property():
    """README.md:85-90: decode_from_code_indices(tokenize(v)) == forward(v, return_recon=True)."""
    g = load_golden("mini")
    model = build_product(g["kwargs"], g["wseed"])
    orc = build_oracle(model, g["kwargs"])
    v = golden_video(g)
    assert torch.equal(orc.decode_from_code_indices(orc.tokenize(v)), orc.forward(v, return_recon=True))


def test_flat_ids_decode_equals_4d():
    """M:1587-1591."""
    g = load_golden("mini")
    model = build_product(g["kwargs"], g["wseed"])
    orc = build_oracle(model, g["kwargs"])
    codes = g["codes"]
    a = orc.decode_from_code_indices(codes)
    b = orc.decode_from_code_indices(codes.reshape(codes.shape[0], -1))
    assert torch.equal(a, b)


def test_temporal_causality_and_batch_independence():
    """SURVEY.md 4 items 2 and 4 on the restated oracle."""
    g = load_golden("mini")
    model = build_product(g["kwargs"], g["wseed"])
    orc = build_oracle(model, g["kwargs"])
    v = golden_video(g)
    c0 = orc.tokenize(v)
    v2 = v.clone()
    v2[:, :, 5:] += 1.0
    c1 = orc.tokenize(v2)
    assert torch.equal(c0[:, :2], c1[:, :2])          # latent frames 0-1 unaffected by frames >= 5
    assert torch.equal(orc.tokenize(v[:1]), c0[:1])    # batch independence


def test_dead_layernorm_never_applied():
    """the final LayerNorm is in state_dict but zip() drops it (M:1565)."""
    g = load_golden("cfg1")
    model = build_product(g["kwargs"], g["wseed"])
    n = len(g["kwargs"]["layers"])
    assert f"encoder_layers.{n}.1.weight" in model.state_dict()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    sd[f"encoder_layers.{n}.1.weight"] += 100.0
    from oracle.restated import OracleTokenizer
    orc = OracleTokenizer(sd, **g["kwargs"])
    assert torch.equal(orc.tokenize(golden_video(g)), g["codes"])
