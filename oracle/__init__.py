"""CPU oracle for the ChronoEdit denoising hot path — TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import anything from here, and there only as the checker or the reported CPU
baseline.  The product package (`chronoedit_b200/`) never imports `oracle` and has no CPU
fallback: it raises when the CUDA library is missing.

Contents
  dit_oracle.py   restatement of transformer_chronoedit.py (DiT per-step forward)
  vae_oracle.py   restatement of the Wan2.1 3D causal VAE encode/decode (wan2pt1.py)
  ref_loader.py   loads the UNMODIFIED reference files from /root/reference (build container only)
  diffusers_shim/ stand-in for the un-vendored diffusers==0.35.2 symbols the reference imports

Pinning status: the reference ships no golden vectors for this path;
the oracle is pinned against outputs of the reference's own code executed in the build
container (tests/golden/make_golden.py -> tests/golden/*.safetensors).  The diffusers
classes themselves are restated from the published 0.35.2 semantics and cross-checked
against the in-tree DiffSynth implementation of the same network.
"""
