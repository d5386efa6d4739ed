"""Restatement of `ChronoEditPipeline.__call__` around the three hot-path objects -- TEST INFRASTRUCTURE ONLY (oracle/__init__.py).

Restates /root/reference/chronoedit_diffusers/pipeline_chronoedit.py for the path the benchmarks use: `prompt_embeds`,
`negative_prompt_embeds` and `image_embeds` passed in (the UMT5 / CLIP encoders are "next" rows), guardrails disabled,
a tensor `image` in [-1, 1]:
    prepare_latents                                   :392-456
    __call__: timesteps, loop with the temporal-reasoning cut, CFG, scheduler.step, latent de-normalisation,
              one or two VAE decodes                  :667-781
It drives ANY objects with the surface the pipeline touches -- the oracle adapters below (functional
oracle on a state dict, CPU or CUDA), the reference's own modules, or the chronoedit_b200 mirrors -- so that on the GPU box,
where the reference file itself is absent, the same loop can be run once with the oracle modules and once with the CUDA path.

Pinning: tests/test_pipeline_cpu.py executes the UNMODIFIED pipeline_chronoedit.py (through oracle/diffusers_shim) with the
reference's own transformer / VAE twin / flow-UniPC scheduler and requires `run_pipeline` with the oracle adapters to
reproduce its output video bit for bit (bf16 configuration, with and without the temporal-reasoning cut); the reference
outputs are stored as tests/golden/pipeline_*.safetensors by tests/golden/make_golden_pipeline.py.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, List, Optional

import torch

from . import dit_oracle as D
from . import unipc_oracle as U
from . import vae_oracle as V

LATENTS_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508,
                0.4134, -0.0715, 0.5517, -0.3632, -0.1922, -0.9497, 0.2503, -0.2921]   # wan2pt1.py:697-714
LATENTS_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743,
               3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253, 2.8251, 1.9160]         # wan2pt1.py:715-732


# ----------------------------------------------------------------------------------------------
# adapters: the oracle restatements behind the surface the pipeline touches
# ----------------------------------------------------------------------------------------------
class OracleTransformer:
    """dit_oracle.dit_forward behind `transformer(hidden_states=, timestep=, encoder_hidden_states=, encoder_hidden_states_image=,
    attention_kwargs=, return_dict=False)[0]`, `.dtype`, `.config.patch_size` (pipeline_chronoedit.py:653, 715-735)."""

    def __init__(self, sd: Dict[str, torch.Tensor], cfg: D.DiTConfig, dtype: torch.dtype):
        self.sd, self.cfg, self.dtype = sd, cfg, dtype
        self.config = SimpleNamespace(patch_size=cfg.patch_size)
        self.calls = 0

    def __call__(self, hidden_states, timestep, encoder_hidden_states, encoder_hidden_states_image=None, attention_kwargs=None,
                 return_dict=True):
        self.calls += 1
        out = D.dit_forward(self.sd, self.cfg, hidden_states, timestep, encoder_hidden_states, encoder_hidden_states_image)
        return (out,) if not return_dict else SimpleNamespace(sample=out)


class _Posterior:
    def __init__(self, mean):
        self.mean = mean

    def mode(self):
        return self.mean


class OracleVAE:
    """vae_oracle behind `vae.encode(x).latent_dist.mode()`, `vae.decode(z, return_dict=False)[0]`, `.config`, `.dtype`,
    `.temperal_downsample` (pipeline_chronoedit.py:185-186, 427-445, 765-781).  diffusers clamps the decoded video to [-1,1]."""

    def __init__(self, sd: Dict[str, torch.Tensor], cfg: V.VAEConfig, dtype: torch.dtype, clamp: bool = True):
        self.sd, self.cfg, self.dtype, self.clamp = sd, cfg, dtype, clamp
        self.temperal_downsample = list(cfg.temperal_downsample)
        self.config = SimpleNamespace(z_dim=cfg.z_dim, latents_mean=LATENTS_MEAN[: cfg.z_dim], latents_std=LATENTS_STD[: cfg.z_dim])

    def encode(self, x):
        return SimpleNamespace(latent_dist=_Posterior(V.vae_encode(self.sd, self.cfg, x.to(self.dtype))))

    def decode(self, z, return_dict=True):
        out = V.vae_decode(self.sd, self.cfg, z.to(self.dtype), clamp=self.clamp)
        return (out,) if not return_dict else SimpleNamespace(sample=out)


class RefVAEAdapter(torch.nn.Module):
    """The reference's in-tree `WanVAE_` (chronoedit/_src/tokenizers/wan2pt1.py:467-581, executed unmodified) behind the
    AutoencoderKLWan surface, with the diffusers deltas of SURVEY.md 8c: no latent mean/std inside (scale = [0, 1]), decode
    clamped to [-1, 1].  Used only by the golden generator / the pin test in the build container."""

    def __init__(self, wan_vae, clamp: bool = True):
        super().__init__()
        self.model = wan_vae
        self.clamp = clamp
        self.temperal_downsample = list(wan_vae.temperal_downsample)
        self.config = SimpleNamespace(z_dim=wan_vae.z_dim, latents_mean=LATENTS_MEAN[: wan_vae.z_dim], latents_std=LATENTS_STD[: wan_vae.z_dim])

    @property
    def dtype(self):
        return next(self.model.parameters()).dtype

    def encode(self, x):
        return SimpleNamespace(latent_dist=_Posterior(self.model.encode(x.to(self.dtype), [0, 1])))

    def decode(self, z, return_dict=True):
        out = self.model.decode(z.to(self.dtype), [0, 1])
        if self.clamp:
            out = out.clamp(-1.0, 1.0)
        return (out,) if not return_dict else SimpleNamespace(sample=out)


class OracleScheduler:
    """unipc_oracle.UniPCOracle behind the scheduler surface the pipeline touches: set_timesteps(n, device=), .timesteps,
    .order, .step(noise, t, latents, return_dict=False)[0], and the `.model_outputs` / `.last_sample` attributes the
    temporal-reasoning cut slices in place (pipeline_chronoedit.py:668-669, 689, 700-709, 739)."""

    order = 1

    def __init__(self, shift: float = 1.0, num_train_timesteps: int = 1000, cuda_semantics: bool = False):
        # constructor `shift` as in FlowUniPCMultistepScheduler.__init__ (fm_solvers_unipc.py:88-128): applied to the training
        # sigmas AND, because the pipeline calls set_timesteps without a shift, once more per schedule (:203-206)
        self.o = U.UniPCOracle(num_train_timesteps=num_train_timesteps, shift=shift, cuda_semantics=cuda_semantics)

    def set_timesteps(self, num_inference_steps, device=None):
        self.o.set_timesteps(num_inference_steps)
        self.timesteps = self.o.timesteps.to(device) if device is not None else self.o.timesteps

    @property
    def model_outputs(self):
        return self.o.model_outputs

    @model_outputs.setter
    def model_outputs(self, v):
        self.o.model_outputs = v

    @property
    def last_sample(self):
        return self.o.last_sample

    @last_sample.setter
    def last_sample(self, v):
        self.o.last_sample = v

    def step(self, model_output, timestep, sample, return_dict=True):
        out = self.o.step(model_output, sample)
        return (out,) if not return_dict else SimpleNamespace(prev_sample=out)


# ----------------------------------------------------------------------------------------------
# the pipeline
# ----------------------------------------------------------------------------------------------
def prepare_latents(vae, image, batch_size, num_channels_latents, height, width, num_frames, dtype, device, generator, latents,
                    vae_scale_factor_temporal, vae_scale_factor_spatial):
    """ChronoEditPipeline.prepare_latents (pipeline_chronoedit.py:392-456); `image` is the preprocessed [B,3,H,W] tensor."""
    num_latent_frames = (num_frames - 1) // vae_scale_factor_temporal + 1
    latent_height, latent_width = height // vae_scale_factor_spatial, width // vae_scale_factor_spatial
    shape = (batch_size, num_channels_latents, num_latent_frames, latent_height, latent_width)
    if latents is None:
        gen_dev = generator.device if generator is not None else torch.device(device)
        latents = torch.randn(shape, generator=generator, device=gen_dev, dtype=dtype).to(device)   # diffusers randn_tensor
    else:
        latents = latents.to(device=device, dtype=dtype)
    image = image.unsqueeze(2)
    video_condition = torch.cat([image, image.new_zeros(image.shape[0], image.shape[1], num_frames - 1, height, width)], dim=2)
    video_condition = video_condition.to(device=device, dtype=dtype)
    latents_mean = torch.tensor(vae.config.latents_mean).view(1, vae.config.z_dim, 1, 1, 1).to(latents.device, latents.dtype)
    latents_std = 1.0 / torch.tensor(vae.config.latents_std).view(1, vae.config.z_dim, 1, 1, 1).to(latents.device, latents.dtype)
    latent_condition = vae.encode(video_condition).latent_dist.mode()
    latent_condition = latent_condition.repeat(batch_size, 1, 1, 1, 1)
    latent_condition = (latent_condition - latents_mean) * latents_std
    mask_lat_size = torch.ones(batch_size, 1, num_frames, latent_height, latent_width)
    mask_lat_size[:, :, list(range(1, num_frames))] = 0
    first_frame_mask = mask_lat_size[:, :, 0:1]
    first_frame_mask = torch.repeat_interleave(first_frame_mask, dim=2, repeats=vae_scale_factor_temporal)
    mask_lat_size = torch.concat([first_frame_mask, mask_lat_size[:, :, 1:, :]], dim=2)
    mask_lat_size = mask_lat_size.view(batch_size, -1, vae_scale_factor_temporal, latent_height, latent_width)
    mask_lat_size = mask_lat_size.transpose(1, 2)
    mask_lat_size = mask_lat_size.to(latent_condition.device)
    return latents, torch.concat([mask_lat_size, latent_condition], dim=1)


@torch.no_grad()
def run_pipeline(transformer, vae, scheduler, image: torch.Tensor, prompt_embeds: torch.Tensor,
                 negative_prompt_embeds: Optional[torch.Tensor], image_embeds: torch.Tensor, height: int, width: int,
                 num_frames: int = 5, num_inference_steps: int = 4, guidance_scale: float = 5.0,
                 generator: Optional[torch.Generator] = None, latents: Optional[torch.Tensor] = None,
                 enable_temporal_reasoning: bool = False, num_temporal_reasoning_steps: int = 0, attention_kwargs=None,
                 device=None, trace: Optional[List] = None) -> torch.Tensor:
    """ChronoEditPipeline.__call__ (pipeline_chronoedit.py:594-812) with embeddings given and output_type="pt"-before-
    postprocess: returns the decoded video tensor [B, 3, T, H, W] (what `video_processor.postprocess_video` receives)."""
    device = torch.device(device) if device is not None else image.device
    vae_scale_factor_temporal = 2 ** sum(vae.temperal_downsample)
    vae_scale_factor_spatial = 2 ** len(vae.temperal_downsample)
    if num_frames % vae_scale_factor_temporal != 1:
        num_frames = num_frames // vae_scale_factor_temporal * vae_scale_factor_temporal + 1
    num_frames = max(num_frames, 1)
    do_cfg = guidance_scale > 1
    batch_size = prompt_embeds.shape[0]
    transformer_dtype = transformer.dtype
    prompt_embeds = prompt_embeds.to(transformer_dtype)
    if negative_prompt_embeds is not None:
        negative_prompt_embeds = negative_prompt_embeds.to(transformer_dtype)
    image_embeds = image_embeds.repeat(batch_size, 1, 1).to(transformer_dtype)

    scheduler.set_timesteps(num_inference_steps, device=device)
    timesteps = scheduler.timesteps
    image = image.to(device, dtype=torch.bfloat16)                      # :673 (video_processor.preprocess output)
    latents, condition = prepare_latents(vae, image, batch_size, vae.config.z_dim, height, width, num_frames, torch.bfloat16, device,
                                         generator, latents, vae_scale_factor_temporal, vae_scale_factor_spatial)
    for i, t in enumerate(timesteps):
        if enable_temporal_reasoning and i == num_temporal_reasoning_steps:   # :700-709
            latents = latents[:, :, [0, -1]]
            condition = condition[:, :, [0, -1]]
            for j in range(len(scheduler.model_outputs)):
                if scheduler.model_outputs[j] is not None:
                    if latents.shape[-3] != scheduler.model_outputs[j].shape[-3]:
                        scheduler.model_outputs[j] = scheduler.model_outputs[j][:, :, [0, -1]]
            if scheduler.last_sample is not None:
                scheduler.last_sample = scheduler.last_sample[:, :, [0, -1]]
        latent_model_input = torch.cat([latents, condition], dim=1).to(transformer_dtype)
        timestep = t.expand(latents.shape[0])
        noise_pred = transformer(hidden_states=latent_model_input, timestep=timestep, encoder_hidden_states=prompt_embeds,
                                 encoder_hidden_states_image=image_embeds, attention_kwargs=attention_kwargs, return_dict=False)[0]
        if do_cfg:
            noise_uncond = transformer(hidden_states=latent_model_input, timestep=timestep, encoder_hidden_states=negative_prompt_embeds,
                                       encoder_hidden_states_image=image_embeds, attention_kwargs=attention_kwargs, return_dict=False)[0]
            noise_pred = noise_uncond + guidance_scale * (noise_pred - noise_uncond)
        latents = scheduler.step(noise_pred, t, latents, return_dict=False)[0]
        if trace is not None:
            trace.append(latents.detach().clone())
    latents = latents.to(vae.dtype)                                       # :765-774
    latents_mean = torch.tensor(vae.config.latents_mean).view(1, vae.config.z_dim, 1, 1, 1).to(latents.device, latents.dtype)
    latents_std = 1.0 / torch.tensor(vae.config.latents_std).view(1, vae.config.z_dim, 1, 1, 1).to(latents.device, latents.dtype)
    latents = latents / latents_std + latents_mean
    if enable_temporal_reasoning and num_temporal_reasoning_steps > 0:    # :776-781
        video_edit = vae.decode(latents[:, :, [0, -1]], return_dict=False)[0]
        video_reason = vae.decode(latents[:, :, :-1], return_dict=False)[0]
        video = torch.cat([video_reason, video_edit[:, :, 1:]], dim=2)
    else:
        video = vae.decode(latents, return_dict=False)[0]
    return video
