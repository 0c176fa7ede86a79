"""IPAdapter wrapper with the reference's surface (ip_adapter/ip_adapter.py:67-158, 289-317):
`.pipe`, `.device`, `.set_ip_adapter()`, `.load_ip_adapter()`, `.get_image_embeds()`, `.set_scale()`.

The CLIP-ViT-H image encoder in front of the projector is `theatergen_b200.clip_vision.TGCLIPVision`:
pass it as `image_encoder=` (any callable with the transformers surface `.image_embeds` / `.hidden_states` works), or
call `get_image_embeds(clip_image_embeds=...)` with precomputed embeddings.
"""
from __future__ import annotations

import os

import torch

from .. import ops
from ..unet import IPAttnProcessorHandle, TGUNet
from .attention_processor import AttnProcessor, IPAttnProcessor
from .resampler import Resampler


class ImageProjModel:
    """LN(Linear(clip_embeds).reshape(-1, T, D))  (ip_adapter/ip_adapter.py:30-47): one GEMM + one LayerNorm."""

    def __init__(self, cross_attention_dim=1024, clip_embeddings_dim=1024, clip_extra_context_tokens=4):
        self.cross_attention_dim = cross_attention_dim
        self.clip_extra_context_tokens = clip_extra_context_tokens
        self.clip_embeddings_dim = clip_embeddings_dim
        self.w = None

    def to(self, *a, **k):
        return self

    def load_state_dict(self, sd, device="cuda"):
        self.w = {k: v.detach().to(device=device, dtype=torch.float16).contiguous() for k, v in sd.items()}
        return self

    def __call__(self, image_embeds):
        if self.w is None:
            raise RuntimeError("ImageProjModel: load_state_dict first")
        e = image_embeds.to(torch.float16).contiguous()
        if not e.is_cuda:
            raise RuntimeError("ImageProjModel runs on CUDA only (no CPU fallback)")
        n, T, D = e.shape[0], self.clip_extra_context_tokens, self.cross_attention_dim
        t = ops.gemm(e, self.w["proj.weight"], self.w["proj.bias"]).view(n * T, D)
        return ops.layernorm(t, self.w["norm.weight"], self.w["norm.bias"]).view(n, T, D)


class _Pipe:
    """Bare holder when the caller has no diffusers pipeline: `.unet`, `.scheduler`, optional `.vae` etc."""

    def __init__(self, unet, scheduler=None, **extra):
        self.unet, self.scheduler = unet, scheduler
        for k, v in extra.items():
            setattr(self, k, v)

    def to(self, device):
        return self


class IPAdapter:
    def __init__(self, sd_pipe, image_encoder_path=None, ip_ckpt=None, device="cuda", num_tokens=4,
                 image_encoder=None, clip_image_processor=None, clip_embeddings_dim=1024):
        self.device = device
        self.image_encoder_path, self.ip_ckpt, self.num_tokens = image_encoder_path, ip_ckpt, num_tokens
        self.pipe = sd_pipe.to(self.device) if hasattr(sd_pipe, "to") else sd_pipe
        self.image_encoder, self.clip_image_processor = image_encoder, clip_image_processor
        self.clip_embeddings_dim = clip_embeddings_dim
        self.set_ip_adapter()
        self.image_proj_model = self.init_proj()
        if ip_ckpt is not None:
            self.load_ip_adapter()

    def init_proj(self):
        return ImageProjModel(cross_attention_dim=self.pipe.unet.config.cross_attention_dim,
                              clip_embeddings_dim=self.clip_embeddings_dim,
                              clip_extra_context_tokens=self.num_tokens)

    def set_ip_adapter(self):
        unet = self.pipe.unet
        if isinstance(unet, TGUNet):
            # module-level drop-in: the handles are already in unet.attn_processors; the context split (text | image
            # tokens, attention_processor.py:467-471) is fixed when the UNet is packed
            if unet.num_ip_tokens != self.num_tokens:
                raise RuntimeError(f"TGUNet was built for {unet.num_ip_tokens} image tokens, the adapter has "
                                   f"{self.num_tokens}: construct TGUNet(state_dict, num_ip_tokens={self.num_tokens})")
            return
        attn_procs = {}
        for name in unet.attn_processors.keys():  # same naming logic as ip_adapter/ip_adapter.py:98-116
            cross = None if name.endswith("attn1.processor") else unet.config.cross_attention_dim
            if name.startswith("mid_block"):
                hidden = unet.config.block_out_channels[-1]
            elif name.startswith("up_blocks"):
                hidden = list(reversed(unet.config.block_out_channels))[int(name[len("up_blocks.")])]
            else:
                hidden = unet.config.block_out_channels[int(name[len("down_blocks.")])]
            if cross is None:
                attn_procs[name] = AttnProcessor()
            else:
                attn_procs[name] = IPAttnProcessor(hidden_size=hidden, cross_attention_dim=cross, scale=1.0,
                                                   num_tokens=self.num_tokens).to(self.device, dtype=torch.float16)
        unet.set_attn_processor(attn_procs)

    def load_ip_adapter(self, state_dict=None):
        if state_dict is None:
            if os.path.splitext(self.ip_ckpt)[-1] == ".safetensors":
                from safetensors import safe_open
                state_dict = {"image_proj": {}, "ip_adapter": {}}
                with safe_open(self.ip_ckpt, framework="pt", device="cpu") as f:
                    for key in f.keys():
                        if key.startswith("image_proj."):
                            state_dict["image_proj"][key.replace("image_proj.", "")] = f.get_tensor(key)
                        elif key.startswith("ip_adapter."):
                            state_dict["ip_adapter"][key.replace("ip_adapter.", "")] = f.get_tensor(key)
            else:
                state_dict = torch.load(self.ip_ckpt, map_location="cpu")
        self.image_proj_model.load_state_dict(state_dict["image_proj"], device=self.device)
        unet = self.pipe.unet
        if isinstance(unet, TGUNet):
            unet.load_ip_weights(state_dict["ip_adapter"])
        else:
            ip_layers = torch.nn.ModuleList(unet.attn_processors.values())
            ip_layers.load_state_dict(state_dict["ip_adapter"])

    @torch.inference_mode()
    def get_image_embeds(self, pil_image=None, clip_image_embeds=None):
        if pil_image is not None:
            if self.image_encoder is None:
                raise RuntimeError("no image encoder bound: pass image_encoder=TGCLIPVision(...) or clip_image_embeds")
            clip_image = self.clip_image_processor(images=pil_image, return_tensors="pt").pixel_values
            clip_image_embeds = self.image_encoder(clip_image.to(self.device, dtype=torch.float16)).image_embeds
        clip_image_embeds = clip_image_embeds.to(self.device, dtype=torch.float16)
        return self.image_proj_model(clip_image_embeds), self.image_proj_model(torch.zeros_like(clip_image_embeds))

    def set_scale(self, scale):
        for p in self.pipe.unet.attn_processors.values():
            if isinstance(p, (IPAttnProcessor, IPAttnProcessorHandle)):
                p.scale = scale


class IPAdapterPlus(IPAdapter):
    """IP-Adapter with fine-grained features: Resampler over the penultimate CLIP hidden states."""

    def __init__(self, *a, num_tokens=16, clip_hidden_size=1280, **k):
        self.clip_hidden_size = clip_hidden_size
        super().__init__(*a, num_tokens=num_tokens, **k)

    def init_proj(self):
        d = self.pipe.unet.config.cross_attention_dim
        return Resampler(dim=d, depth=4, dim_head=64, heads=12, num_queries=self.num_tokens,
                         embedding_dim=self.clip_hidden_size, output_dim=d, ff_mult=4)

    @torch.inference_mode()
    def get_image_embeds(self, pil_image=None, clip_image_embeds=None, uncond_clip_image_embeds=None):
        """clip_image_embeds: penultimate hidden states [n,257,1280] (hidden_states[-2], ip_adapter.py:310)."""
        if pil_image is not None:
            if self.image_encoder is None:
                raise RuntimeError("no image encoder bound: pass image_encoder=TGCLIPVision(...) or clip_image_embeds")
            clip_image = self.clip_image_processor(images=pil_image, return_tensors="pt").pixel_values
            clip_image = clip_image.to(self.device, dtype=torch.float16)
            clip_image_embeds = self.image_encoder(clip_image, output_hidden_states=True).hidden_states[-2]
            uncond_clip_image_embeds = self.image_encoder(torch.zeros_like(clip_image),
                                                          output_hidden_states=True).hidden_states[-2]
        return self.image_proj_model(clip_image_embeds.to(self.device)), \
            self.image_proj_model(uncond_clip_image_embeds.to(self.device))
