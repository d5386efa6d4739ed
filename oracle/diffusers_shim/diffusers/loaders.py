"""diffusers.loaders names the reference imports (transformer_chronoedit.py:25, pipeline_chronoedit.py:27).

`WanLoraLoaderMixin` restates the call chain of diffusers 0.35.2 ([diffusers-mem]): `pipe.load_lora_weights(x)` resolves a state
dict and hands the `transformer.`-prefixed part to `transformer.load_lora_adapter(...)`; `pipe.fuse_lora(lora_scale=s)` calls
`transformer.fuse_lora(s, ...)`; `unload_lora_weights()` -> `transformer.unload_lora()`.  (run_inference_diffusers.py:369-376.)"""


class FromOriginalModelMixin:
    pass


class PeftAdapterMixin:
    pass


class WanLoraLoaderMixin:
    _lora_loadable_modules = ["transformer"]
    transformer_name = "transformer"

    @classmethod
    def lora_state_dict(cls, pretrained_model_name_or_path_or_dict, **kwargs):
        if isinstance(pretrained_model_name_or_path_or_dict, dict):
            return dict(pretrained_model_name_or_path_or_dict)
        import os

        from safetensors.torch import load_file

        path = pretrained_model_name_or_path_or_dict
        if os.path.isdir(path):
            path = os.path.join(path, kwargs.get("weight_name") or "pytorch_lora_weights.safetensors")
        return load_file(path)

    def load_lora_weights(self, pretrained_model_name_or_path_or_dict, adapter_name=None, **kwargs):
        state_dict = self.lora_state_dict(pretrained_model_name_or_path_or_dict, **kwargs)
        getattr(self, self.transformer_name).load_lora_adapter(state_dict, prefix="transformer", adapter_name=adapter_name, _pipeline=self)

    def fuse_lora(self, components=("transformer",), lora_scale=1.0, safe_fusing=False, adapter_names=None, **kwargs):
        for c in components:
            model = getattr(self, c, None)
            if model is not None:
                model.fuse_lora(lora_scale, safe_fusing=safe_fusing, adapter_names=adapter_names)

    def unload_lora_weights(self):
        for c in self._lora_loadable_modules:
            model = getattr(self, c, None)
            if model is not None and hasattr(model, "unload_lora"):
                model.unload_lora()
