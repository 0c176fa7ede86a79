"""Multi-GPU partitioning of the hot path: one process per GPU, independent units, no per-timestep collective.

Units: dialogues are fully independent (generate.py:183-188).  Inside a dialogue the only cross-unit
state is the reference image / IP tokens of an `obj_id` written by its first stage-1 run (models/pipelines.py:183-199,
476-477), which gives a three-phase DAG per dialogue:
    A  first-appearance stage-1 runs          (IP scale 0, no dependency)
    B  stage-1 runs of already-known ids      (need A's IP tokens, scale 0.4)
    C  stage-2 run of every turn              (needs that turn's stage-1 latents + masks)
Collectives: one broadcast of the packed UNet/IP weights at start-up, one tiny all-gather of IP tokens between
A and B (only when a dialogue's characters are split over ranks).  Seeds are a pure function of (dialogue index, turn,
repeat) exactly as generate.py:157-243 computes them in a fresh run -- never of execution order.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch
import torch.distributed as dist

LARGE_CONSTANT = 123456789   # generate.py:157
LARGE_CONSTANT3 = 6789       # generate.py:159


def turn_seeds(dialogue_index, turn_index, repeat_ind=0, seed_offset=0, n_dialogues=None, regenerate_ind=0,
               freeze_dialogue_seed=None, turns_per_dialogue=4, turns_before=None):
    """(bg_seed, fg_seed_start) of generate.py:209-243 for a fresh run (no resume skips).  `seed_offset` grows by one
    per finished dialogue (generate.py:262).  `ind` (the seed base of the `--freeze_dialogue_seed None` mode) counts
    the turns that EXIST: generate.py:195-198 `continue`s past a missing turn before `ind += 1` (:260) is reached.
    `turns_before` = number of existing turns generated before this one (all earlier dialogues + earlier turns of this
    dialogue); the default assumes `turns_per_dialogue` turns in every earlier dialogue and none missing before
    `turn_index` -- true for every CMIGBench dialogue (4000 x 4 turns)."""
    done_dialogues = dialogue_index + (regenerate_ind * n_dialogues if regenerate_ind else 0)
    ind = done_dialogues * turns_per_dialogue + turn_index if turns_before is None else turns_before
    base = freeze_dialogue_seed if freeze_dialogue_seed is not None else ind
    ind_offset = repeat_ind * LARGE_CONSTANT3 + seed_offset + done_dialogues
    bg_seed = base + ind_offset
    return bg_seed, bg_seed + LARGE_CONSTANT


def count_turns(dialogue):
    """turns generate.py's loop finds in a dialogue record (`for turn in ['turn 1'..'turn 4']`, generate.py:191-198)"""
    return sum(1 for t in range(4) if f"turn {t + 1}" in dialogue)


def shard_dialogues(names, world, rank):
    """Round-robin over ranks (config 5: 64 dialogues over 8 GPUs = 8 each, zero traffic after the broadcast)."""
    return [n for i, n in enumerate(names) if i % world == rank]


@dataclass
class WorkItem:
    phase: str            # "A" | "B" | "C" | "alias"
    dialogue: str
    turn: int             # 0-based
    obj_index: int = -1   # position inside the turn's object list (-1 for phase C)
    obj_id: int = -1
    phrase: str = ""
    box_xywh: tuple = ()
    ip_scale: float = 0.0
    bg_seed: int = 0
    fg_seed_start: int = 0
    deps: list = field(default_factory=list)
    alias_of: int = -1    # phase "alias": obj_index (same turn) of the item whose outputs this position repeats


def build_dag(dialogue_name, dialogue, dialogue_index=0, repeat_ind=0, seed_offset=0, freeze_dialogue_seed=0,
              turns_before_dialogue=None):
    """dialogue: {"turn 1": {"objects": [[phrase, [x,y,w,h], obj_id], ...], ...}, ...} (CMIGBench story.json).
    `freeze_dialogue_seed` defaults to 0 like generate.py's command line (generate.py:45, 209-212: every turn of a
    dialogue shares the base seed); None selects the turn-counter seeds of the `--freeze_dialogue_seed None` mode
    (`turns_before_dialogue` = existing turns of all earlier dialogues; default 4 per dialogue).
    Returns {"A": [...], "B": [...], "C": [...], "alias": [...], "turns": {t: [item per object, in order]}}.

    Duplicates follow theatergen.py:217-226 exactly: an object is skipped only when the IMMEDIATELY PREVIOUS object of
    the turn has the same (prompt, obj_id) -- its slot then repeats the previous slot's outputs (an "alias" item, no
    work).  A non-consecutive repeat is generated again; by then the id's reference image exists (the sequential
    reference wrote it at the first occurrence, models/pipelines.py:476-477), so it is a phase-B item at scale 0.4.
    The per-object prompt is a function of the phrase alone inside a turn (parse.convert_spec), so (phrase, obj_id)
    stands for (prompt, obj_id).  Every work item of an id -- phase A included -- is keyed by (turn, obj_index)."""
    phases = {"A": [], "B": [], "C": [], "alias": [], "turns": {}}
    seen = set()
    n_before = (dialogue_index * 4 if turns_before_dialogue is None else turns_before_dialogue)
    for t in range(4):
        turn = dialogue.get(f"turn {t + 1}")
        if turn is None:
            continue
        bg, fg = turn_seeds(dialogue_index, t, repeat_ind, seed_offset, freeze_dialogue_seed=freeze_dialogue_seed,
                            turns_before=n_before)
        n_before += 1
        prev, prev_index, slots = None, -1, []
        for oi, (phrase, box, obj_id) in enumerate(turn["objects"]):
            if prev == (phrase, obj_id):
                it = WorkItem("alias", dialogue_name, t, oi, obj_id, phrase, tuple(box), bg_seed=bg, fg_seed_start=fg,
                              alias_of=prev_index)
                phases["alias"].append(it)
                slots.append(it)
                continue
            prev, prev_index = (phrase, obj_id), oi
            first = obj_id not in seen
            seen.add(obj_id)
            it = WorkItem("A" if first else "B", dialogue_name, t, oi, obj_id, phrase, tuple(box),
                          0.0 if first else 0.4, bg, fg, deps=[] if first else [("A", obj_id)])
            phases[it.phase].append(it)
            slots.append(it)
        phases["turns"][t] = slots
        phases["C"].append(WorkItem("C", dialogue_name, t, ip_scale=0.1, bg_seed=bg, fg_seed_start=fg,
                                    deps=[("stage1", t)]))
    return phases


def assign(items, world):
    """Work items of one DAG phase -> per-rank lists, item i on rank i % world (config 4)."""
    out = [[] for _ in range(world)]
    for i, it in enumerate(items):
        out[i % world].append(it)
    return out


def cobatch(items, max_chars=4):
    """Group a rank's stage-1 items into co-batched forwards (B = 2 x len(group), per-sample IP scale)."""
    return [items[i:i + max_chars] for i in range(0, len(items), max_chars)]


# ------------------------------------------------------------------------------------------------------------------
# collectives (NCCL on GPUs, gloo in the CPU tests)
# ------------------------------------------------------------------------------------------------------------------
def broadcast_tensors(tensors, src=0, group=None, bucket_bytes=256 << 20):
    """Broadcast a list of same-dtype tensors in a few large flat buckets (launch latency, not link count, is what
    matters on NVSwitch)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    bucket, size = [], 0
    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([t.reshape(-1) for t in bucket])
        dist.broadcast(flat, src, group=group)
        off = 0
        for t in bucket:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()
        bucket, size = [], 0
    for t in tensors:
        bucket.append(t)
        size += t.numel() * t.element_size()
        if size >= bucket_bytes:
            flush()
    flush()


def unet_weight_tensors(unet):
    """Every packed weight tensor of a TGUNet in a deterministic order."""
    out = [unet.w[k] for k in sorted(unet.w)]
    for group in (unet.resnets, unet.transformers):
        for name in sorted(group):
            out += [group[name][k] for k in sorted(group[name]) if torch.is_tensor(group[name][k])]
    return out


def broadcast_unet_weights(unet, src=0, group=None):
    broadcast_tensors(unet_weight_tensors(unet), src, group)


def all_gather_ip_tokens(local_ids, local_tokens, group=None):
    """Exchange between phases A and B: {obj_id: tokens[2,T,768]} known on every rank afterwards.
    local_ids: list[int]; local_tokens: tensor [len(local_ids), 2, T, D] (cond, uncond)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return {i: local_tokens[k] for k, i in enumerate(local_ids)}
    world = dist.get_world_size(group)
    payload = (list(local_ids), local_tokens.cpu() if local_tokens is not None else None)
    gathered = [None] * world
    dist.all_gather_object(gathered, payload, group=group)  # a few KB per id: latency only
    out = {}
    for ids, toks in gathered:
        for k, i in enumerate(ids):
            out[i] = toks[k].to(local_tokens.device) if local_tokens is not None else toks[k]
    return out
