"""DagRunner -- executes the work-item DAG of a CMIGBench dialogue (sharding.build_dag) on the DenoiseEngine, one
process per GPU (BASELINE.json configs[3]: "CMIGBench story task, 4-turn dialogue, 4 characters, shard turns over
4xB200").

What the reference does sequentially on one GPU (generate.py:183-247 -> theatergen.run :278-488 ->
get_masked_latents_all_list :204-273 -> pipelines.generate_semantic_guidance per character -> compose ->
pipelines.final_image_generation per turn) becomes three phases with one small exchange between them:

    A  stage 1 of every first appearance of an obj_id          IP scale 0 (placeholder image, pipelines.py:195-199)
         -> the id's reference image -> image tokens            (the reference writes database/{obj_id}.png :476-477 and
                                                                 re-reads + re-encodes it on every later turn :183-204;
                                                                 here the tokens are cached per id: IPTokenCache)
       all-gather of the new ids' tokens [2, T, 768] fp16       (a few KB per id)
    B  stage 1 of every later appearance                        IP scale 0.4, tokens of the id's reference image
       all-gather of the stage-1 traces [steps+1, 4, 64, 64] fp16 + 64x64 masks (1.7 MB per character)
    C  stage 2 of every turn on the rank that owns it            align + compose (utils/latents.py) -> ControlNet + UNet
                                                                 (IP scale 0.1, tokens of the turn's first character)
                                                                 + masked composite on every step

Inside a phase a rank runs its items in co-batched forwards of EXACTLY `cobatch` trajectories (the last group is padded
by repeating an item): a trajectory's bits depend on the batch size of the forward it shares (tile / split-K selection per
GEMM shape) but never on its position or its neighbours, so with a fixed group size every trajectory is bit-identical
whatever the world size and the assignment -- tests/test_runner_gpu.py checks 1 rank against 2 ranks bit for bit.

The models on either side of the hot path -- text encoder, CLIP image encoder + VAE decode of the reference image, SAM,
the line-art detector -- are reached through a small `models` object.  `SyntheticModels` provides seeded stand-ins with
the right shapes for the benchmark and the tests (no weights exist offline): prompt embeddings seeded by the prompt's
hash, "CLIP" embeddings as a fixed function of the character's final latents (so phase B really depends on phase A's
bits), the centred layout box as the SAM mask, mask edges as the line-art hint.  Real encoders plug in with the same four
methods.
"""
from __future__ import annotations

import hashlib
import time
from dataclasses import dataclass, field

import torch
import torch.distributed as dist

from . import sharding
from .utils import latents as L
from .utils import utils as U

NEGATIVE_SO = "background, multiple objects, incomplete, lowres, bad anatomy, low quality, obscured"  # pipelines.py:221


class SyntheticModels:
    """Seeded stand-ins for the models outside the hot path (see the module docstring)."""

    def __init__(self, device, num_tokens=4):
        self.device, self.num_tokens = torch.device(device), num_tokens

    def encode_prompt(self, prompt, negative_prompt):
        """-> (positive [1,77,768], negative [1,77,768]) fp16 (adapter.pipe.encode_prompt, pipelines.py:238-245)"""
        def emb(s):
            g = torch.Generator().manual_seed(int(hashlib.md5(s.encode()).hexdigest()[:12], 16))
            return torch.randn(1, 77, 768, generator=g).half().to(self.device)
        return emb("p:" + str(prompt)), emb("n:" + str(negative_prompt))

    def clip_image_embeds(self, final_latents):
        """[1,4,64,64] fp16 -> [1,1024] fp16: stand-in for VAE decode -> PNG -> CLIP-ViT-H image_embeds
        (pipelines.py:468-477, ip_adapter.py:142-153): 4x4 average pooling of the latents, a fixed function of their bits"""
        p = torch.nn.functional.avg_pool2d(final_latents.float(), 4).reshape(1, 1024)
        return (p / p.norm().clamp_min(1e-6) * 32.0).half()

    def placeholder_image_embeds(self):
        """the embedding of model.png, used at scale 0 for a first appearance (pipelines.py:195-199)"""
        g = torch.Generator().manual_seed(20240428)
        return torch.randn(1, 1024, generator=g).half().to(self.device)

    def object_mask(self, so_box, final_latents=None):
        """stand-in for GroundingDINO + SAM (theatergen.py:141-182): the centred layout box as a 64x64 bool mask"""
        return U.proportion_to_mask(so_box, 64, 64, device=self.device).bool()

    def control_hint(self, union_mask):
        """stand-in for the line-art detector on the pasted collage (pipelines.py:710-731): the mask's edges, [1,3,512,512]"""
        m = union_mask.float()[None, None]
        edge = (torch.nn.functional.max_pool2d(m, 3, 1, 1) - m).clamp(0, 1)
        return torch.nn.functional.interpolate(edge, size=(512, 512), mode="nearest").repeat(1, 3, 1, 1).half()


class IPTokenCache:
    """obj_id -> image tokens [2, T, 768] fp16 (row 0 conditional, row 1 unconditional): what the reference recomputes
    from database/{obj_id}.png on every later turn (pipelines.py:183-213).  `save` / `load` keep the cache next to the
    reference's character database so a resumed run (generate.py:192-194 skips finished turns) finds it."""

    def __init__(self):
        self.tokens = {}

    def __contains__(self, obj_id):
        return obj_id in self.tokens

    def put(self, obj_id, cond, uncond):
        self.tokens[obj_id] = torch.cat([cond, uncond]).to(torch.float16).contiguous()

    def get(self, obj_id):
        t = self.tokens[obj_id]
        return t[0:1], t[1:2]

    def save(self, path):
        torch.save({k: v.cpu() for k, v in self.tokens.items()}, path)

    def load(self, path, device):
        self.tokens.update({k: v.to(device) for k, v in torch.load(path, map_location="cpu").items()})
        return self


@dataclass
class RunStats:
    phase_ms: dict = field(default_factory=dict)        # phase -> wall time of the phase on this rank (incl. waiting)
    busy_ms: dict = field(default_factory=dict)         # phase -> time this rank spent denoising
    denoise_steps: dict = field(default_factory=dict)   # phase -> useful denoise steps of this rank (padding excluded)
    padded_steps: dict = field(default_factory=dict)    # phase -> steps spent on padding trajectories


def box_xywh_to_norm(box, size=512):
    """CMIGBench [x, y, w, h] @512 -> normalised xyxy (utils/parse.py convert_spec)"""
    x, y, w, h = box
    return [x / size, y / size, (x + w) / size, (y + h) / size]


def centered_box(box_norm):
    """stage-1 objects are generated centred (theatergen.py:348-350: horizontal + floor padding 0.2)"""
    return U.get_centered_box(box_norm, horizontal_center_only=False, vertical_placement="floor_padding",
                              floor_padding=0.2)


class DagRunner:
    def __init__(self, engine, adapter, models=None, cobatch=2, num_steps=50, guidance_scale=7.5, group=None,
                 task="story"):
        self.eng, self.adapter, self.group = engine, adapter, group
        self.models = models or SyntheticModels(engine.device, adapter.num_tokens)
        self.K, self.steps, self.guidance = int(cobatch), int(num_steps), float(guidance_scale)
        self.task = task
        self.dev = engine.device
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    # ------------------------------------------------------------------------------------------------------------
    def _stage1_inputs(self, item, cache):
        """initial latents (CPU-seeded fp16 blend, utils/latents.py:257-295), context [uncond; cond] x (77 + T), IP scale"""
        box = box_xywh_to_norm(item.box_xywh)
        so_box = centered_box(box)
        lat, _, _ = L.get_input_latents_list(None, item.bg_seed, item.fg_seed_start, 0.01, 512, 512, self.adapter,
                                             so_boxes=[so_box])
        prompt = ("single object, " if self.task == "editing" else "full-body picture of ") + item.phrase
        pos, neg = self.models.encode_prompt(prompt, NEGATIVE_SO)
        if item.phase == "A":   # no reference yet: placeholder image at scale 0 (the branch is computed and multiplied by 0)
            cond, uncond = self.adapter.get_image_embeds(clip_image_embeds=self.models.placeholder_image_embeds())
        else:
            cond, uncond = cache.get(item.obj_id)
        return dict(item=item, latents=lat[0].to(self.dev, torch.float16), so_box=so_box, box=box,
                    ctx_u=torch.cat([neg, uncond], 1), ctx_c=torch.cat([pos, cond], 1), scale=float(item.ip_scale))

    def _run_stage1(self, jobs, stats, phase):
        """jobs: list of _stage1_inputs dicts of THIS rank -> {(turn, obj_index): (trace [S+1,1,4,64,64], mask, final)}"""
        out = {}
        K, S = self.K, self.steps
        busy = 0.0
        for g0 in range(0, len(jobs), K):
            grp = jobs[g0:g0 + K]
            n_real = len(grp)
            grp = grp + [grp[-1]] * (K - n_real)         # fixed batch size: see the module docstring
            lat = torch.cat([j["latents"] for j in grp])
            ctx = torch.cat([j["ctx_u"] for j in grp] + [j["ctx_c"] for j in grp]).half().contiguous()
            scale = torch.tensor([j["scale"] for j in grp] * 2, dtype=torch.float32, device=self.dev)
            trace = torch.empty((S + 1, K, 4, 64, 64), dtype=torch.float16, device=self.dev)
            trace[0].copy_(lat)
            t0 = time.perf_counter()
            final = self.eng.denoise(lat, ctx, scale, num_steps=S, guidance_scale=self.guidance,
                                     on_step=lambda i, sg: trace[i + 1].copy_(sg.trace, non_blocking=True))
            torch.cuda.synchronize(self.dev)
            busy += time.perf_counter() - t0
            for k in range(n_real):
                j = grp[k]
                fin = final[k:k + 1].clone()
                mask = self.models.object_mask(j["so_box"], fin)
                out[(j["item"].turn, j["item"].obj_index)] = (trace[:, k:k + 1].contiguous(), mask, fin)
            stats.denoise_steps[phase] = stats.denoise_steps.get(phase, 0) + n_real * S
            stats.padded_steps[phase] = stats.padded_steps.get(phase, 0) + (K - n_real) * S
        stats.busy_ms[phase] = stats.busy_ms.get(phase, 0.0) + busy * 1e3
        return out

    # ------------------------------------------------------------------------------------------------------------
    def _gather_tokens(self, local):
        """{obj_id: tokens[2,T,768]} of this rank -> the union over ranks (sharding.all_gather_ip_tokens)"""
        ids = sorted(local)
        toks = torch.stack([local[i] for i in ids]) if ids else torch.zeros((0, 2, self.adapter.num_tokens, 768),
                                                                           dtype=torch.float16, device=self.dev)
        return sharding.all_gather_ip_tokens(ids, toks, self.group)

    def _gather_stage1(self, local, keys_by_rank):
        """Every rank receives every stage-1 result of the dialogue: one padded all_gather of the traces (fp16) and one of
        the masks over NCCL.  keys_by_rank[r] = ordered (turn, obj_index) keys produced by rank r."""
        if self.world == 1:
            return dict(local)
        S = self.steps
        n_max = max(1, max(len(k) for k in keys_by_rank))
        mine = keys_by_rank[self.rank]
        tr = torch.zeros((n_max, S + 1, 1, 4, 64, 64), dtype=torch.float16, device=self.dev)
        mk = torch.zeros((n_max, 64, 64), dtype=torch.uint8, device=self.dev)
        for i, k in enumerate(mine):
            tr[i].copy_(local[k][0])
            mk[i].copy_(local[k][1].to(torch.uint8))
        if dist.get_backend(self.group) == "nccl":
            tr_all = torch.empty((self.world, *tr.shape), dtype=tr.dtype, device=self.dev)
            mk_all = torch.empty((self.world, *mk.shape), dtype=mk.dtype, device=self.dev)
            dist.all_gather_into_tensor(tr_all, tr, group=self.group)
            dist.all_gather_into_tensor(mk_all, mk, group=self.group)
        else:  # gloo (the CPU / single-GPU tests): staged through host memory
            trs = [torch.empty(tr.shape, dtype=tr.dtype) for _ in range(self.world)]
            mks = [torch.empty(mk.shape, dtype=mk.dtype) for _ in range(self.world)]
            dist.all_gather(trs, tr.cpu(), group=self.group)
            dist.all_gather(mks, mk.cpu(), group=self.group)
            tr_all, mk_all = torch.stack(trs).to(self.dev), torch.stack(mks).to(self.dev)
        out = {}
        for r, keys in enumerate(keys_by_rank):
            for i, k in enumerate(keys):
                out[k] = (tr_all[r, i], mk_all[r, i].bool(), tr_all[r, i, -1])
        return out

    # ------------------------------------------------------------------------------------------------------------
    def _stage2_inputs(self, citem, slots, stage1, cache):
        """One turn: align the characters' stage-1 traces to their layout boxes and compose them (utils/latents.py:168-255;
        the reference computes this and then overwrites it, SURVEY.md finding 6 -- kept for the foreground index map),
        then the tensors final_image_generation builds (pipelines.py:604-632, 700-738): the frozen trajectory = the pasted
        composite re-noised to every timestep with the turn's CUDA generator, the background noise, the layout mask, the
        ControlNet hint, the context with the FIRST character's image tokens (finding 8)."""
        S = self.steps
        res = []
        for it in slots:
            key = (it.turn, it.alias_of if it.phase == "alias" else it.obj_index)
            res.append(stage1[key])
        traces = [r[0] for r in res]
        masks = [r[1] for r in res]
        boxes = [box_xywh_to_norm(it.box_xywh) for it in slots]
        bg = L.get_unscaled_latents(1, 4, 512, 512, torch.manual_seed(citem.bg_seed), torch.float16)
        composed, fg_idx, _, _ = L.compose_latents_with_alignment(
            "1.5", self.adapter, 0, None, None, None, traces, masks, S, 1, 512, 512, latents_bg=bg,
            align_with_overall_bboxes=True, overall_bboxes=[[b] for b in boxes])
        union = fg_idx != 0                                                    # frozen_mask of theatergen.py:442
        gen = torch.Generator(self.dev).manual_seed(citem.bg_seed)           # pipelines.py:594
        x0 = composed[S]                                                       # stand-in for VAE-encode(pasted collage)
        noise = torch.randn(x0.shape, generator=gen, device=self.dev, dtype=torch.float16)
        my_bg = torch.randn(x0.shape, generator=gen, device=self.dev, dtype=torch.float16)
        sched = self.eng.scheduler
        sched.set_timesteps(S)
        frozen = torch.cat([my_bg[None], sched.add_noise(x0, noise, sched.timesteps).unsqueeze(1)]).half()
        overall = ", ".join(it.phrase for it in slots)
        pos, neg = self.models.encode_prompt(overall, "incohesive, edge shadow, blurry")
        cond, uncond = cache.get(slots[0].obj_id)
        return dict(item=citem, latents=my_bg, frozen=frozen.contiguous(), mask=union.to(torch.float16),
                    ctx=torch.cat([torch.cat([neg, uncond], 1), torch.cat([pos, cond], 1)]).half().contiguous(),
                    ctrl_ctx=torch.cat([neg, pos]).half().contiguous(), hint=self.models.control_hint(union),
                    fg_idx=fg_idx)

    def _run_stage2(self, jobs, stats):
        """Stage-2 trajectories run one per forward (B = 2 + the B = 2 ControlNet branch): each has its own mask."""
        out = {}
        busy = 0.0
        S = self.steps
        for j in jobs:
            cn = self.eng.controlnet
            kw = {}
            if cn is not None:
                kw = dict(control_ctx=j["ctrl_ctx"], cond_emb=cn.embed_condition(j["hint"]))
            t0 = time.perf_counter()
            final = self.eng.denoise(j["latents"], j["ctx"], 0.1, num_steps=S, guidance_scale=self.guidance,
                                     frozen_all=j["frozen"], frozen_mask=j["mask"], frozen_steps=S, **kw)
            torch.cuda.synchronize(self.dev)
            busy += time.perf_counter() - t0
            out[j["item"].turn] = (final.clone(), j["fg_idx"])
            stats.denoise_steps["C"] = stats.denoise_steps.get("C", 0) + S
        stats.busy_ms["C"] = stats.busy_ms.get("C", 0.0) + busy * 1e3
        return out

    # ------------------------------------------------------------------------------------------------------------
    def _barrier(self):
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            dist.barrier(group=self.group)

    @torch.no_grad()
    def run_dialogue(self, name, dialogue, dialogue_index=0, cache=None, **dag_kwargs):
        """-> (stage1 {(turn, obj_index): final latents}, stage2 {turn: (final latents, fg index map)} of the turns this rank
        owns, RunStats).  Every rank ends up with every stage-1 result; stage-2 results stay on the owning rank."""
        dag = sharding.build_dag(name, dialogue, dialogue_index, **dag_kwargs)
        cache = cache if cache is not None else IPTokenCache()
        stats = RunStats()
        results = {}
        produced = {}   # rank -> ordered keys over phases A and B

        def phase_stage1(phase):
            per_rank = sharding.assign(dag[phase], self.world)
            for r in range(self.world):
                produced.setdefault(r, []).extend((it.turn, it.obj_index) for it in per_rank[r])
            self._barrier()
            t0 = time.perf_counter()
            jobs = [self._stage1_inputs(it, cache) for it in per_rank[self.rank]]
            local = self._run_stage1(jobs, stats, phase)
            results.update(local)
            self._barrier()
            stats.phase_ms[phase] = (time.perf_counter() - t0) * 1e3
            return per_rank[self.rank], local

        # ---- phase A: first appearances; their reference image becomes the id's image tokens ----
        mine, local = phase_stage1("A")
        new_tokens = {}
        for it in mine:
            fin = local[(it.turn, it.obj_index)][2]
            cond, uncond = self.adapter.get_image_embeds(clip_image_embeds=self.models.clip_image_embeds(fin))
            new_tokens[it.obj_id] = torch.cat([cond, uncond]).half()
        for obj_id, tok in self._gather_tokens(new_tokens).items():
            cache.put(obj_id, tok[0:1], tok[1:2])
        # ---- phase B: later appearances (scale 0.4) ----
        phase_stage1("B")
        # ---- exchange: every stage-1 trace / mask to every rank ----
        t0 = time.perf_counter()
        stage1 = self._gather_stage1(results, [produced.get(r, []) for r in range(self.world)])
        self._barrier()
        stats.phase_ms["exchange"] = (time.perf_counter() - t0) * 1e3
        # ---- phase C: stage 2 per turn ----
        per_rank = sharding.assign(dag["C"], self.world)
        t0 = time.perf_counter()
        jobs = [self._stage2_inputs(c, dag["turns"][c.turn], stage1, cache) for c in per_rank[self.rank]]
        stage2 = self._run_stage2(jobs, stats)
        self._barrier()
        stats.phase_ms["C"] = (time.perf_counter() - t0) * 1e3
        return {k: v[2] for k, v in stage1.items()}, stage2, stats
