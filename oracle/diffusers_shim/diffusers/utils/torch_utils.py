"""diffusers.utils.torch_utils.randn_tensor ([diffusers-mem] 0.35.2): draw on the generator's device, then move."""
import torch


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    layout = layout or torch.strided
    device = torch.device(device) if device is not None else torch.device("cpu")
    rand_device = device
    if generator is not None:
        gen = generator[0] if isinstance(generator, list) else generator
        if gen.device.type != device.type and gen.device.type == "cpu":
            rand_device = torch.device("cpu")
    if isinstance(generator, list):
        shape = (1,) + tuple(shape[1:])
        latents = [torch.randn(shape, generator=generator[i], device=rand_device, dtype=dtype, layout=layout) for i in range(len(generator))]
        return torch.cat(latents, dim=0).to(device)
    return torch.randn(shape, generator=generator, device=rand_device, dtype=dtype, layout=layout).to(device)
