"""diffusers.image_processor: `PipelineImageInput` (annotation only, pipeline_chronoedit.py:26) and the slice of
VaeImageProcessor.preprocess / postprocess that VideoProcessor uses ([diffusers-mem] 0.35.2): PIL / tensor -> [B,3,H,W] in [-1,1]."""
from typing import List, Union

import numpy as np
import PIL.Image
import torch
import torch.nn.functional as F

PipelineImageInput = Union[PIL.Image.Image, np.ndarray, torch.Tensor, List[PIL.Image.Image], List[np.ndarray], List[torch.Tensor]]


class VaeImageProcessor:
    def __init__(self, do_resize=True, vae_scale_factor=8, resample="lanczos", do_normalize=True):
        self.do_resize, self.vae_scale_factor, self.resample, self.do_normalize = do_resize, vae_scale_factor, resample, do_normalize

    @staticmethod
    def normalize(x):
        return 2.0 * x - 1.0

    @staticmethod
    def denormalize(x):
        return (x * 0.5 + 0.5).clamp(0, 1)

    def preprocess(self, image, height=None, width=None):
        if isinstance(image, PIL.Image.Image):
            image = [image]
        if isinstance(image, list) and isinstance(image[0], PIL.Image.Image):
            if self.do_resize and height is not None:
                image = [im.resize((width, height), resample=PIL.Image.Resampling.LANCZOS) for im in image]
            arr = np.stack([np.array(im.convert("RGB")).astype(np.float32) / 255.0 for im in image], axis=0)
            t = torch.from_numpy(arr.transpose(0, 3, 1, 2))
        else:
            t = torch.cat(image, dim=0) if isinstance(image, list) and image[0].ndim == 4 else (torch.stack(image, dim=0) if isinstance(image, list) else image)
            if t.ndim == 3:
                t = t.unsqueeze(0)
            if self.do_resize and height is not None and tuple(t.shape[-2:]) != (height, width):
                t = F.interpolate(t, size=(height, width))
        if self.do_normalize and not (isinstance(image, torch.Tensor) and t.min() < 0):   # tensors already in [-1,1] are passed through
            t = self.normalize(t)
        return t

    def postprocess(self, image, output_type="pil"):
        if output_type in ("latent", "pt"):
            return image if output_type == "latent" else self.denormalize(image)
        image = self.denormalize(image)
        arr = image.cpu().permute(0, 2, 3, 1).float().numpy()
        if output_type == "np":
            return arr
        return [PIL.Image.fromarray((a * 255).round().astype("uint8")) for a in arr]
