import torch.nn as nn


class ModelMixin(nn.Module):
    _keep_in_fp32_modules = None

    @property
    def dtype(self):
        """diffusers.models.modeling_utils.get_parameter_dtype ([diffusers-mem] 0.35.2): dtype of the first floating-point
        parameter that is not one of `_keep_in_fp32_modules` (the fp32 scale_shift_table must not make a bf16 model "fp32")."""
        keep = self._keep_in_fp32_modules or []
        last = None
        for name, p in self.named_parameters():
            last = p.dtype
            if any(k in name for k in keep):
                continue
            if p.is_floating_point():
                return p.dtype
        return last

    @property
    def device(self):
        return next(self.parameters()).device
