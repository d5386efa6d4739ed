"""CPU tests of the VAE checkpoint surface: the complete diffusers <-> twin parameter-name map, `load_state_dict` on a
diffusers-layout state dict, and `AutoencoderKLWan.from_pretrained(dir, subfolder="vae")` on a synthetic diffusers-layout
checkpoint (run_inference_diffusers.py:341-345) -- every tensor must land on the parameter the oracle VAE reads it from."""
import json
import os

import pytest
import torch
from safetensors.torch import save_file

from oracle import vae_oracle as V


def _twin_weights(cfg):
    return {k: v.to(torch.bfloat16) for k, v in V.random_state_dict(cfg, seed=3).items()}


@pytest.mark.parametrize("cfg", [V.VAEConfig.wan21(), V.VAEConfig.tiny(32)], ids=["wan21", "tiny"])
def test_diffusers_key_map_is_a_bijection_onto_the_oracle_parameters(cfg):
    from chronoedit_b200.autoencoder import AutoencoderKLWan

    kmap = AutoencoderKLWan.diffusers_key_map(cfg.dim, cfg.z_dim, tuple(cfg.dim_mult), cfg.num_res_blocks, tuple(cfg.temperal_downsample))
    twin = V.param_shapes(cfg)
    assert sorted(kmap.values()) == sorted(twin), "every parameter of the oracle VAE must be the image of exactly one diffusers key"
    assert len(kmap) == len(twin)
    # spot checks of the published diffusers module tree
    assert kmap["encoder.down_blocks.0.conv1.weight"] == "encoder.downsamples.0.residual.2.weight"
    assert kmap["encoder.down_blocks.2.resample.1.weight"] == "encoder.downsamples.2.resample.1.weight"
    assert kmap["decoder.up_blocks.0.upsamplers.0.time_conv.weight"] == "decoder.upsamples.3.time_conv.weight"
    assert kmap["decoder.up_blocks.1.resnets.0.conv_shortcut.weight"] == "decoder.upsamples.4.shortcut.weight"
    assert kmap["decoder.up_blocks.3.resnets.2.conv2.bias"] == "decoder.upsamples.14.residual.6.bias"
    assert kmap["decoder.mid_block.attentions.0.to_qkv.weight"] == "decoder.middle.1.to_qkv.weight"
    assert kmap["post_quant_conv.weight"] == "conv2.weight" and kmap["quant_conv.bias"] == "conv1.bias"


def test_from_pretrained_round_trips_a_diffusers_layout_checkpoint(tmp_path):
    from chronoedit_b200 import CEError
    from chronoedit_b200.autoencoder import AutoencoderKLWan

    cfg = V.VAEConfig.tiny(32)
    twin = _twin_weights(cfg)
    kmap = AutoencoderKLWan.diffusers_key_map(cfg.dim, cfg.z_dim, tuple(cfg.dim_mult), cfg.num_res_blocks, tuple(cfg.temperal_downsample))
    dif = {d: twin[t].contiguous() for d, t in kmap.items()}
    vae_dir = tmp_path / "ckpt" / "vae"
    os.makedirs(vae_dir)
    with open(vae_dir / "config.json", "w") as f:
        json.dump({"_class_name": "AutoencoderKLWan", "base_dim": cfg.dim, "z_dim": cfg.z_dim, "dim_mult": list(cfg.dim_mult),
                   "num_res_blocks": cfg.num_res_blocks, "attn_scales": [], "temperal_downsample": list(cfg.temperal_downsample), "dropout": 0.0}, f)
    save_file(dif, str(vae_dir / "diffusion_pytorch_model.safetensors"))
    m = AutoencoderKLWan.from_pretrained(str(tmp_path / "ckpt"), subfolder="vae", torch_dtype=torch.bfloat16)
    own = dict(m.named_parameters())
    assert sorted(own) == sorted(twin)
    for k, v in twin.items():
        assert torch.equal(own[k].data, v), k
    assert m.config.z_dim == cfg.z_dim and m.temperal_downsample == list(cfg.temperal_downsample)
    # and back
    back = m.diffusers_state_dict()
    assert sorted(back) == sorted(dif) and all(torch.equal(back[k], dif[k]) for k in dif)
    # anything that does not map is an error, not a silent skip
    bad = dict(dif)
    bad["decoder.up_blocks.0.attentions.0.norm.gamma"] = torch.zeros(4)
    with pytest.raises(CEError, match="unmapped"):
        AutoencoderKLWan(base_dim=cfg.dim, z_dim=cfg.z_dim, dim_mult=tuple(cfg.dim_mult), num_res_blocks=cfg.num_res_blocks).load_state_dict(bad)
    short = dict(dif)
    short.pop("encoder.mid_block.resnets.1.conv2.weight")
    with pytest.raises(CEError, match="missing"):
        AutoencoderKLWan(base_dim=cfg.dim, z_dim=cfg.z_dim, dim_mult=tuple(cfg.dim_mult), num_res_blocks=cfg.num_res_blocks).load_state_dict(short)


def test_twin_names_still_load():
    from chronoedit_b200.autoencoder import AutoencoderKLWan

    cfg = V.VAEConfig.tiny(32)
    twin = _twin_weights(cfg)
    m = AutoencoderKLWan(base_dim=cfg.dim, z_dim=cfg.z_dim, dim_mult=tuple(cfg.dim_mult), num_res_blocks=cfg.num_res_blocks)
    m.load_state_dict(twin, strict=True)
    assert all(torch.equal(p.data, twin[k]) for k, p in m.named_parameters())
