"""Mirror of the reference's models/pipelines.py for the SD-1.5 path: same function names, positional signatures and
return tuples (generate_semantic_guidance :175-490, final_image_generation :592-857, prepare_ip_embeds :860-950,
latent_backward_guidance :62-128, encode/decode :130-173), with the per-timestep hot loops executed by
theatergen_b200.engine.DenoiseEngine (CUDA-graph replayed libtheatergen.so kernels).

Reached through the objects the caller passes in, exactly like the reference does (duck-typed):
`adapter.pipe.encode_prompt` (its `text_encoder` can be theatergen_b200.clip_text.TGCLIPText), `adapter.pipe.vae` /
`image_processor` (theatergen_b200.vae.TGAutoencoderKL), `adapter.get_image_embeds(pil_image=...)`
(theatergen_b200.clip_vision.TGCLIPVision), `processor` (the line-art detector: not re-implemented).  `controlnetpipe.controlnet` is a
theatergen_b200.controlnet.TGControlNet on the kernel path (any other module is called eagerly per step).  `basever == 'xl'` raises: SDXL is outside the metric.

Reference quirks kept on purpose: the stage-1 prompt is re-encoded from `ip_prompt` and the
`input_embeddings` argument is ignored (#3); stage 2 overwrites `latents_all[0]`, `latents_all[1:]` and `frozen_mask`
(#6/#9); the composite uses `latents_all[index+1]` (#5); IP scale 0.4 / 0 in stage 1 and 0.1 in stage 2 (#2);
stage-2 image tokens come from the first character only (#8); ControlNet conditioning scale is always 1.0 (#6).
"""
from __future__ import annotations

import warnings

import numpy as np
import torch

from .. import ops
from ..controlnet import TGControlNet
from ..engine import DenoiseEngine
from ..unet import TGUNet
from ..utils import guidance, schedule

torch_device = "cuda"
MAX_GUIDANCE_GRAPHS = 4  # captured guidance-gradient graphs kept per UNet (least recently used is dropped)
DEFAULT_GUIDANCE_ATTN_KEYS = [("mid", 0, 0, 0), ("up", 1, 0, 0), ("up", 1, 1, 0), ("up", 1, 2, 0)]  # pipelines.py:21
NEGATIVE_SO = "background, multiple objects, incomplete, lowres, bad anatomy, low quality, obscured"

_ENGINES = {}


def _engine(adapter) -> DenoiseEngine:
    unet = adapter.pipe.unet
    if not isinstance(unet, TGUNet):
        raise RuntimeError("theatergen_b200.models.pipelines needs adapter.pipe.unet to be a TGUNet "
                           "(module-level drop-in); there is no eager/PyTorch fallback")
    key = id(unet)
    if key not in _ENGINES:
        _ENGINES[key] = DenoiseEngine(unet, adapter.pipe.scheduler)
    return _ENGINES[key]


def get_unscaled_latents(batch_size, in_channels, height, width, generator, dtype):
    return torch.randn((batch_size, in_channels, height // 8, width // 8), generator=generator, dtype=dtype,
                       device=torch_device)


def get_scaled_latents(batch_size, in_channels, height, width, generator, dtype, scheduler):
    return get_unscaled_latents(batch_size, in_channels, height, width, generator, dtype) * scheduler.init_noise_sigma


def _open_reference_image(database_path, obj_id):
    from PIL import Image
    try:
        return Image.open(database_path + str(obj_id) + ".png"), 0.4, 1
    except Exception:
        return Image.open("model.png"), 0, 0  # placeholder, scale 0 (pipelines.py:195-199; opened relative to CWD)


def latent_backward_guidance(adapter, scheduler, unet, cond_embeddings, index, bboxes, object_positions, t, latents,
                             loss, loss_scale=30, loss_threshold=0.2, max_iter=5, max_index_step=10,
                             cross_attention_kwargs=None, ref_ca_saved_attns=None, guidance_attn_keys=None,
                             verbose=False, clear_cache=False, prompt_embeds=None, final=False, **kwargs):
    """pipelines.py:62-128 on our kernels: while the energy is above the threshold, run the conditional UNet forward with
    in-kernel capture of the guidance maps, the fused energy kernel, and d(energy * loss_scale)/d(latents) through the
    recorded forward (theatergen_b200/backward.py -- the reference uses autograd, :102), then
    `latents -= grad * sqrt(1 - alpha_prod_t)` (:106-115, the DDIM branch: the scheduler has no `sigmas`).
    Returns (latents, loss) like the reference; `loss` is a 0-d float32 CUDA tensor (energy * loss_scale)."""
    from ..backward import GuidanceBackward, GuidanceStepGraph
    iteration = 0
    if index < max_index_step:
        if isinstance(max_iter, list):
            max_iter = max_iter[index] if len(max_iter) > index else max_iter[-1]
        if not isinstance(unet, TGUNet):
            raise RuntimeError("latent_backward_guidance needs the kernel-path UNet (TGUNet)")
        gb = getattr(unet, "_guidance_backward", None)
        if gb is None:
            gb = unet._guidance_backward = GuidanceBackward(unet)
        keys = guidance_attn_keys if guidance_attn_keys is not None else DEFAULT_GUIDANCE_ATTN_KEYS
        owned = False
        bad = torch.zeros(1, dtype=torch.int32, device=latents.device)
        while loss.item() / loss_scale > loss_threshold and iteration < max_iter and index < max_index_step:
            if len(bboxes) == 0:
                break
            if ref_ca_saved_attns is None and getattr(unet, "guidance_graphs", True):
                # forward + energy + backward replayed as one CUDA graph per (layout, prompt) -- built on first use
                # (the layout is baked into the graph by value; the prompt embedding is a static buffer refreshed below)
                gkey = repr((bboxes, object_positions, keys, sorted(kwargs.items()), loss_scale, tuple(latents.shape),
                             tuple(cond_embeddings.shape), gb._ip_scale()))
                graphs = unet.__dict__.setdefault("_guidance_graphs", {})
                sg = graphs.pop(gkey, None)
                if sg is None:
                    while len(graphs) >= MAX_GUIDANCE_GRAPHS:  # each graph pins its activations (about 2 GB): evict the oldest
                        graphs.pop(next(iter(graphs)))
                    plan, _ = guidance.build_plan(bboxes, object_positions, keys, index=index, **kwargs)
                    sg = GuidanceStepGraph(gb, plan, cond_embeddings, loss_scale, tuple(latents.shape))
                    sg.capture(scheduler.scale_model_input(latents, t), t)
                graphs[gkey] = sg  # most recently used last
                sg.ctx.copy_(cond_embeddings)
                loss, grad = sg.replay(scheduler.scale_model_input(latents, t), t)
            else:
                plan, ref_maps = guidance.build_plan(bboxes, object_positions, keys, ref_ca_saved_attns=ref_ca_saved_attns,
                                                     index=index, **kwargs)
                loss, grad = gb.energy_and_grad(scheduler.scale_model_input(latents, t), t, cond_embeddings, plan,
                                                loss_scale=loss_scale, ref_maps=ref_maps)
            if torch.isnan(loss):
                print("**Loss is NaN**")
            if not owned:  # the reference rebinds `latents` (out of place); the caller's tensor stays untouched
                latents, owned = latents.to(torch.float16).clone(), True
            alpha_prod_t = float(scheduler.alphas_cumprod[int(t)])
            ops.guidance_update(latents, grad, 1.0 / gb.grad_scale, (1.0 - alpha_prod_t) ** 0.5, bad=bad)
            iteration += 1
            if verbose:
                print(f"time index {index}, loss: {loss.item() / loss_scale:.3f} (de-scaled with scale {loss_scale:.1f}), "
                      f"iteration {iteration}")
        if int(bad):
            warnings.warn("latent_backward_guidance: non-finite gradient values were skipped (fp16 overflow)")
    return latents, loss


def encode(model_dict, image, generator):
    """pipelines.py:130-160 (VAE encode: outside the hot path, uses the caller's VAE)."""
    vae, dtype = model_dict.vae, model_dict.dtype
    if isinstance(image, np.ndarray) or not torch.is_tensor(image):
        image = torch.from_numpy(np.array(image).astype(np.float32) / 255.0)[None].permute(0, 3, 1, 2)
        image = 2.0 * image - 1.0
    image = image.to(device=torch_device, dtype=dtype)
    return vae.config.scaling_factor * vae.encode(image).latent_dist.sample(generator)


def decode(vae, latents):
    """pipelines.py:162-173"""
    with torch.no_grad():
        image = vae.decode(1 / 0.18215 * latents).sample
    image = (image / 2 + 0.5).clamp(0, 1)
    image = image.detach().cpu().permute(0, 2, 3, 1).numpy()
    return (image * 255).round().astype("uint8")


@torch.no_grad()
def generate_semantic_guidance(task, fg_seed_now, basever, ip_prompt, database_path, id, adapter, model_dict, latents,
                               input_embeddings, num_inference_steps, bboxes, phrases, object_positions,
                               guidance_scale=7.5, semantic_guidance_kwargs=None, return_cross_attn=False,
                               return_saved_cross_attn=False, saved_cross_attn_keys=None, return_cond_ca_only=False,
                               return_token_ca_only=None, offload_guidance_cross_attn_to_cpu=False,
                               offload_cross_attn_to_cpu=False, offload_latents_to_cpu=True, return_box_vis=False,
                               show_progress=True, save_all_latents=False, dynamic_num_inference_steps=False,
                               fast_after_steps=None, fast_rate=2, use_boxdiff=False, use_adapter=False, obj_id=False,
                               have_reffer=0, clip_image_embeds=None):
    """Stage 1: one character, 50 x (UNet B=2 with IP tokens -> CFG -> DDIM), all latents traced.
    `clip_image_embeds` (extra keyword) lets a caller without a CLIP vision encoder supply the image embedding."""
    if basever == 'xl':
        raise NotImplementedError("SDXL is outside the hot path this build replaces")
    eng = _engine(adapter)
    if obj_id:
        if clip_image_embeds is None:
            pil_image, scale, have_reffer = _open_reference_image(database_path, obj_id)
        else:
            pil_image, scale = None, (0.4 if have_reffer else 0)
        adapter.set_scale(scale)
        image_prompt_embeds, uncond_image_prompt_embeds = adapter.get_image_embeds(
            pil_image=pil_image, clip_image_embeds=clip_image_embeds)
        prompt = ("single object, " if task == 'editing' else "full-body picture of ") + str(ip_prompt)
        with torch.inference_mode():
            prompt_embeds_, negative_prompt_embeds_ = adapter.pipe.encode_prompt(
                prompt, device=adapter.device, num_images_per_prompt=1, do_classifier_free_guidance=True,
                negative_prompt=NEGATIVE_SO)
        prompt_embeds = torch.cat([prompt_embeds_, image_prompt_embeds], dim=1)
        negative_prompt_embeds = torch.cat([negative_prompt_embeds_, uncond_image_prompt_embeds], dim=1)
    else:
        raise ValueError("generate_semantic_guidance: obj_id is required (the reference leaves prompt_embeds undefined)")
    prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds]).half()  # uncond first (pipelines.py:370)

    scheduler = adapter.pipe.scheduler
    scheduler.set_timesteps(num_inference_steps)
    if fast_after_steps is not None:
        scheduler.timesteps = schedule.get_fast_schedule(scheduler.timesteps, fast_after_steps, fast_rate)
        raise NotImplementedError("fast schedule is inactive in the reference configuration (use_fast_schedule=False)")
    latents = latents.clone().to(torch_device, torch.float16)
    n = num_inference_steps
    trace = torch.empty((n + 1, 1, 4, 64, 64), dtype=torch.float16).pin_memory() if save_all_latents else None
    scale = next(p.scale for p in adapter.pipe.unet.attn_processors.values() if hasattr(p, "scale"))
    saved_attns = [{} for _ in range(n)] if return_saved_cross_attn else None  # cross_attention_kwargs=None upstream
    final = eng.denoise(latents, prompt_embeds, float(scale), num_steps=n, guidance_scale=guidance_scale,
                        trace_host=trace)
    latents = final.clone().half()
    torch.cuda.current_stream().synchronize()  # the pinned trace is complete from here on

    image = adapter.pipe.vae.decode(latents / adapter.pipe.vae.config.scaling_factor, return_dict=False)[0].detach()
    images = adapter.pipe.image_processor.postprocess(image, output_type='pil', do_denormalize=[True])[0]
    if have_reffer == 0:
        images.save(database_path + str(obj_id) + ".png")  # first appearance -> character database
    ret = [latents, images]
    if return_cross_attn:
        ret.append(([], [], []))
    if return_saved_cross_attn:
        ret.append(saved_attns)
    if return_box_vis:
        ret.append(images)
    if save_all_latents:
        ret.append(trace if offload_latents_to_cpu else trace.to(torch_device))
    return tuple(ret)


def generate(model_dict, latents, input_embeddings, num_inference_steps, guidance_scale=7.5, no_set_timesteps=False,
             scheduler_key='dpm_scheduler', adapter=None):
    """pipelines.py:492-521: plain CFG text-to-image loop (text-only context)."""
    text_embeddings, uncond_embeddings, cond_embeddings = input_embeddings
    eng = _engine(adapter)
    final = eng.denoise(latents.to(torch_device, torch.float16), text_embeddings.half(), 0.0,
                        num_steps=num_inference_steps, guidance_scale=guidance_scale)
    return final.clone(), decode(model_dict.vae, final)


def prepare_ip_embeds(height, width, generator, device, basever, adapter, bg_prompt, overall_negative_prompt,
                      single_obj_img_list, objects, clip_image_embeds=None):
    """pipelines.py:860-950, SD-1.5 branch: [negative ; positive] x (77 text + T image tokens)."""
    if basever == 'xl':
        raise NotImplementedError("SDXL is outside the hot path this build replaces")
    image_prompt_embeds, uncond_image_prompt_embeds = adapter.get_image_embeds(
        pil_image=single_obj_img_list if clip_image_embeds is None else None, clip_image_embeds=clip_image_embeds)
    with torch.inference_mode():
        prompt_embeds_, negative_prompt_embeds_ = adapter.pipe.encode_prompt(
            bg_prompt, device=adapter.device, num_images_per_prompt=1, do_classifier_free_guidance=True,
            negative_prompt=overall_negative_prompt)
    prompt_embeds = torch.cat([prompt_embeds_, image_prompt_embeds], dim=1)
    negative_prompt_embeds = torch.cat([negative_prompt_embeds_, uncond_image_prompt_embeds], dim=1)
    return torch.cat([negative_prompt_embeds, prompt_embeds]), None, None


@torch.no_grad()
def final_image_generation(basever, processor, controlnetpipe, tpipe, overall_prompt, overall_negative_prompt,
                           bg_prompt, single_obj_img_list, objects, repeat_ind, height, width, bg_seed, inp_mask,
                           input_img, adapter, model_dict, latents_all, frozen_mask, bg_input_embeddings,
                           input_embeddings, num_inference_steps, frozen_steps, guidance_scale=7.5, bboxes=None,
                           phrases=None, object_positions=None, semantic_guidance_kwargs=None,
                           offload_guidance_cross_attn_to_cpu=False, use_boxdiff=False, clip_image_embeds=None,
                           stage2_noise=None):
    """Stage 2: 50 x (ControlNet -> UNet(+residuals, IP scale 0.1) -> CFG -> DDIM -> masked composite).
    `stage2_noise=(init_latents, noise, bg_latents)` lets tests inject the tensors the reference draws from a CUDA
    generator (not reproducible on the host)."""
    if basever == 'xl':
        raise NotImplementedError("SDXL is outside the hot path this build replaces")
    vae, unet, scheduler, dtype = adapter.pipe.vae, adapter.pipe.unet, adapter.pipe.scheduler, adapter.pipe.unet.dtype
    eng = _engine(adapter)
    generator = torch.Generator("cuda").manual_seed(bg_seed)
    scheduler.set_timesteps(num_inference_steps)

    # mask: 1 - (resize(inp_mask, 64x64) > 0)   (pipelines.py:604-614)
    m = inp_mask.resize((int(height / 8), int(width / 8))).convert('L')
    m = np.array(m).astype(np.float32) / 255.0
    m[m > 0] = 1
    my_mask = torch.from_numpy(1 - m).to(device=torch_device, dtype=dtype)

    if stage2_noise is None:
        img = np.array(input_img).astype(np.float32) / 255.0
        img = torch.from_numpy(img[None].transpose(0, 3, 1, 2))
        myimage = (2.0 * img - 1.0).to(device=torch_device, dtype=dtype)
        init_latents = vae.config.scaling_factor * vae.encode(myimage).latent_dist.sample(generator=generator)
        noise = torch.randn(init_latents.shape, generator=generator, device=torch_device, dtype=dtype)
        my_bg = get_scaled_latents(1, unet.config.in_channels, height, width, generator, dtype, scheduler)
    else:
        init_latents, noise, my_bg = [t.to(torch_device, dtype) for t in stage2_noise]
    my_latents = scheduler.add_noise(init_latents, noise, scheduler.timesteps).unsqueeze(1)  # [T,1,4,64,64]

    ip_embeds, _, _ = prepare_ip_embeds(height, width, generator, torch_device, basever, adapter, overall_prompt,
                                        overall_negative_prompt, single_obj_img_list[0], objects,
                                        clip_image_embeds=clip_image_embeds)
    adapter.set_scale(0.1)

    latents = my_bg
    frozen_mask = my_mask
    frozen = torch.cat([latents[None].to(dtype), my_latents.to(dtype)], dim=0)  # latents_all[0]=bg, [1:]=noised collage
    if torch.is_tensor(latents_all) and latents_all.shape == frozen.shape:
        latents_all.copy_(frozen)  # the reference mutates the caller's tensor in place (:737-738)

    controlnet = getattr(controlnetpipe, "controlnet", None) if controlnetpipe is not None else None
    control_ctx = cond_emb = None
    if isinstance(controlnet, TGControlNet):
        # ControlNet on our kernels, fused into the UNet forward and graph-replayed (pipelines.py:702-731, 770-778).
        # The conditioning embedder is timestep independent: it runs once here instead of once per step.
        control_image = processor(np.array(input_img))
        imagetight = controlnetpipe.prepare_image(image=control_image, width=width, height=height, batch_size=1,
                                                  num_images_per_prompt=1, device="cuda", dtype=torch.float16,
                                                  do_classifier_free_guidance=True, guess_mode=False)
        cond_emb = controlnet.embed_condition(imagetight[:1])  # the CFG pair shares one hint image
        control_ctx = input_embeddings[0].to(torch_device, torch.float16)  # [uncond ; cond] x 77 text tokens (:761)
        eng.controlnet = controlnet
    elif controlnet is not None:
        # a foreign ControlNet module (e.g. stock diffusers): one eager forward per step feeds NCHW residuals to TGUNet
        return _stage2_with_controlnet(eng, adapter, controlnetpipe, processor, input_img, input_embeddings, ip_embeds,
                                       latents, frozen, frozen_mask, num_inference_steps, frozen_steps, guidance_scale,
                                       height, width, vae)
    n = num_inference_steps
    trace = torch.empty((n + 1, 1, 4, 64, 64), dtype=torch.float16).pin_memory()
    final = eng.denoise(latents.half(), ip_embeds.half(), 0.1, num_steps=n, guidance_scale=guidance_scale,
                        frozen_all=frozen.half(), frozen_mask=frozen_mask, frozen_steps=frozen_steps, trace_host=trace,
                        control_ctx=control_ctx, cond_emb=cond_emb)
    latents = final.clone()
    torch.cuda.current_stream().synchronize()
    images = decode(vae, latents)
    return latents, images


def _stage2_with_controlnet(eng, adapter, controlnetpipe, processor, input_img, input_embeddings, ip_embeds, latents,
                            frozen, frozen_mask, n, frozen_steps, guidance_scale, height, width, vae):
    from .. import ops
    text_embeddings = input_embeddings[0]
    controlnet = controlnetpipe.controlnet
    control_image = processor(np.array(input_img))
    imagetight = controlnetpipe.prepare_image(image=control_image, width=width, height=height, batch_size=1,
                                              num_images_per_prompt=1, device="cuda", dtype=controlnet.dtype,
                                              do_classifier_free_guidance=True, guess_mode=False)
    sched = eng.scheduler
    ts, coef = eng.tables(n, guidance_scale)
    latents = latents.half().contiguous().clone()
    unet = eng.unet
    for index, t in enumerate(sched.timesteps):
        x_in = torch.cat([latents] * 2)
        down, mid = controlnet(x_in, t, encoder_hidden_states=text_embeddings, controlnet_cond=imagetight,
                               conditioning_scale=1.0, guess_mode=False, return_dict=False)
        eps = unet(x_in, t, encoder_hidden_states=ip_embeds, down_block_additional_residuals=down,
                   mid_block_additional_residual=mid)[0]
        comp = index < frozen_steps
        ops.cfg_ddim_composite(eps, latents, coef[index], frozen=frozen[index + 1].half().contiguous() if comp else None,
                               mask=frozen_mask.half().contiguous() if comp else None)
    return latents, decode(vae, latents)
