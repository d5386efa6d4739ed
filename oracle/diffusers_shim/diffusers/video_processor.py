"""diffusers.video_processor.VideoProcessor ([diffusers-mem] 0.35.2): preprocess of the conditioning image happens through the
image path (pipeline_chronoedit.py:673); postprocess_video turns [B,C,T,H,W] in [-1,1] into np [B,T,H,W,C] in [0,1] / pt / pil (:802)."""
import numpy as np
import torch

from .image_processor import VaeImageProcessor


class VideoProcessor(VaeImageProcessor):
    def postprocess_video(self, video, output_type="np"):
        outputs = []
        for b in range(video.shape[0]):
            outputs.append(self.postprocess(video[b].permute(1, 0, 2, 3), output_type))
        if output_type == "np":
            return np.stack(outputs)
        if output_type == "pt":
            return torch.stack(outputs)
        return outputs
