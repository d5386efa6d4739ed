"""Host side of the encoder mirrors against `transformers` (the package the reference pipeline imports its encoders from,
pipeline_chronoedit.py:23): parameter names / shapes of a checkpoint, the T5 relative-position bucket function, and the
no-CPU-fallback rule.  The arithmetic is compared on the GPU (tests/test_gpu_encoders.py)."""
import pytest
import torch

import chronoedit_b200 as ce
from chronoedit_b200._lib import CEError

transformers = pytest.importorskip("transformers")

UMT5_KW = dict(vocab_size=512, d_model=64, d_kv=16, d_ff=128, num_layers=2, num_heads=4)
CLIP_KW = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, image_size=28, patch_size=14)


def test_umt5_state_dict_contract_and_bucket_function():
    hf = transformers.UMT5EncoderModel(transformers.UMT5Config(**UMT5_KW, feed_forward_proj="gated-gelu", relative_attention_num_buckets=32,
                                                               relative_attention_max_distance=128))
    ours = ce.UMT5EncoderModel(**UMT5_KW)
    hf_sd = hf.state_dict()
    mine = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    theirs = {k: tuple(v.shape) for k, v in hf_sd.items() if k != "encoder.embed_tokens.weight"}   # tied to shared.weight
    assert mine == theirs
    assert torch.equal(hf_sd["encoder.embed_tokens.weight"], hf_sd["shared.weight"])
    missing, unexpected = ours.load_state_dict(hf_sd, strict=True)   # a transformers checkpoint loads as it is
    assert not missing and not unexpected
    # the bucket function behind the per-layer bias tables (bidirectional T5 buckets), every key - query distance up to 2048
    att = hf.encoder.block[0].layer[0].SelfAttention
    rel = torch.arange(-2048, 2049)
    assert torch.equal(ours._bucket(rel), att._relative_position_bucket(rel))
    with pytest.raises(CEError, match="no CPU path"):
        ours(torch.zeros(1, 8, dtype=torch.long))


def test_clip_vision_state_dict_contract():
    hf = transformers.CLIPVisionModel(transformers.CLIPVisionConfig(**CLIP_KW, hidden_act="quick_gelu"))
    ours = ce.CLIPVisionModel(**CLIP_KW, hidden_act="quick_gelu")
    mine = {k: tuple(v.shape) for k, v in ours.state_dict().items()}
    theirs = {k: tuple(v.shape) for k, v in hf.state_dict().items() if not k.endswith("position_ids")}
    assert mine == theirs
    missing, unexpected = ours.load_state_dict(hf.state_dict(), strict=True)
    assert not missing and not unexpected
    with pytest.raises(CEError, match="hidden_states"):          # the pooled head is deliberately not built (pipeline_chronoedit.py:258)
        ours(torch.zeros(1, 3, 28, 28))
    with pytest.raises(CEError, match="no CPU path"):
        ours(torch.zeros(1, 3, 28, 28), output_hidden_states=True).hidden_states[-2]
