"""CPU: the parameter tables in gcd_b200/spec.py equal the reference modules' state_dict (golden key dumps made by
oracle/pin_against_reference.py from the reference's own VideoUNet / VideoDecoder)."""
import json
import os

import pytest

from gcd_b200 import spec

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("fname,shapes", [
    ("unet_tiny_keys.json", lambda: spec.unet_param_shapes(spec.UNET_TINY)),
    ("unet_kubric_keys.json", lambda: spec.unet_param_shapes(spec.UNET_KUBRIC)),
    ("unet_pardom_keys.json", lambda: spec.unet_param_shapes(spec.UNET_PARDOM)),
    ("vae_tiny_keys.json", lambda: spec.decoder_param_shapes(spec.VAE_TINY)),
    ("vae_full_keys.json", lambda: spec.decoder_param_shapes(spec.VAE_DECODER)),
    ("enc_tiny_keys.json", lambda: spec.encoder_param_shapes(spec.VAE_ENCODER_TINY)),
    ("enc_full_keys.json", lambda: spec.encoder_param_shapes(spec.VAE_ENCODER)),
])
def test_param_tables_match_reference(fname, shapes):
    path = os.path.join(GOLD, fname)
    if not os.path.exists(path):
        pytest.skip(f"{fname} not generated")
    ref = {k: tuple(v) for k, v in json.load(open(path)).items()}
    mine = {k: tuple(v) for k, v in shapes().items()}
    assert set(mine) == set(ref), (sorted(set(ref) - set(mine))[:5], sorted(set(mine) - set(ref))[:5])
    bad = [k for k in ref if ref[k] != mine[k]]
    assert not bad, [(k, ref[k], mine[k]) for k in bad[:5]]


def test_module_state_dict_keys():
    import torch  # noqa: F401
    from gcd_b200.unet import VideoUNet
    net = VideoUNet(**spec.unet_ctor_kwargs(spec.UNET_TINY))
    assert list(net.state_dict().keys()) == list(spec.unet_param_shapes(spec.UNET_TINY).keys())
    n = sum(p.numel() for p in net.parameters())
    assert n > 1e6
    from gcd_b200.vae import Encoder
    enc = Encoder(**spec.encoder_ctor_kwargs(spec.VAE_ENCODER))
    assert list(enc.state_dict().keys()) == list(spec.encoder_param_shapes(spec.VAE_ENCODER).keys())
    assert sum(p.numel() for p in enc.parameters()) == 34_163_592      # 34 M parameters


def test_kubric_param_count():
    n = 0
    for shp in spec.unet_param_shapes(spec.UNET_KUBRIC).values():
        k = 1
        for s in shp:
            k *= s
        n += k
    assert n == 1_526_427_882  # SURVEY.md §6 [probe]
    n = 0
    for shp in spec.decoder_param_shapes(spec.VAE_DECODER).values():
        k = 1
        for s in shp:
            k *= s
        n += k
    assert n == 63_579_183
