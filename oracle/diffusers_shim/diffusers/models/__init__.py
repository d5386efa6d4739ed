"""`from diffusers.models import AutoencoderKLWan, WanTransformer3DModel` (pipeline_chronoedit.py:28): annotations only in that
file; the golden generator wraps the reference's in-tree VAE twin in oracle/pipeline_oracle.py instead."""


class AutoencoderKLWan:   # placeholder for the annotation at pipeline_chronoedit.py:167
    pass


class WanTransformer3DModel:
    pass
