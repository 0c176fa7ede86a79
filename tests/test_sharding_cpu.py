"""Host logic of the multi-GPU path: DAG, seeds, assignment -- and the collectives on a world_size-2 gloo group."""
import json
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from theatergen_b200 import sharding as S

DIALOGUE4 = {  # CMIGBench story "dialogue 4"-shaped fixture: 2+2+2+4 character generations, 4 ids
    "turn 1": {"objects": [["a girl", [10, 20, 100, 200], 0], ["a dog", [200, 220, 120, 150], 1]]},
    "turn 2": {"objects": [["a girl", [30, 20, 100, 200], 0], ["a dog", [220, 240, 120, 150], 1]]},
    "turn 3": {"objects": [["a girl", [30, 20, 100, 200], 0], ["a cat", [300, 300, 100, 100], 2]]},
    "turn 4": {"objects": [["a girl", [278, 251, 157, 185], 0], ["a dog", [358, 328, 150, 180], 1],
                           ["a cat", [168, 4, 100, 80], 2], ["a bird", [4, 331, 80, 70], 3]]},
}


def test_seed_arithmetic_matches_generate_py():
    # generate.py: ind counts turns globally, seed_offset += 1 per dialogue, bg = ind + repeat*6789 + seed_offset
    assert S.turn_seeds(0, 0) == (0, 123456789)
    assert S.turn_seeds(0, 3) == (3, 3 + 123456789)
    assert S.turn_seeds(2, 1) == (2 * 4 + 1 + 2, 2 * 4 + 1 + 2 + 123456789)
    assert S.turn_seeds(2, 1, repeat_ind=1, seed_offset=5)[0] == 9 + 6789 + 5 + 2
    assert S.turn_seeds(5, 2, freeze_dialogue_seed=7)[0] == 7 + 5


def test_dag_phases_and_assignment():
    dag = S.build_dag("dialogue 4", DIALOGUE4, dialogue_index=3)
    assert [len(dag[p]) for p in "ABC"] == [4, 6, 4]                      # 10 character generations, 4 turns
    assert sorted(i.obj_id for i in dag["A"]) == [0, 1, 2, 3]
    assert all(i.ip_scale == 0.0 for i in dag["A"]) and all(i.ip_scale == 0.4 for i in dag["B"])
    assert all(i.ip_scale == 0.1 for i in dag["C"])
    assert all(("A", i.obj_id) in i.deps for i in dag["B"])
    per_rank = S.assign(dag["B"], 4)
    assert [len(r) for r in per_rank] == [2, 2, 1, 1] and sum(len(r) for r in per_rank) == 6
    assert [len(g) for g in S.cobatch(dag["B"], 4)] == [4, 2]
    names = [f"dialogue {i}" for i in range(64)]
    shards = [S.shard_dialogues(names, 8, r) for r in range(8)]
    assert all(len(s) == 8 for s in shards) and sorted(sum(shards, [])) == sorted(names)
    # seeds depend on (dialogue index, turn) only -- not on which rank runs the item
    assert dag["C"][2].bg_seed == S.turn_seeds(3, 2, freeze_dialogue_seed=0)[0] == 3   # generate.py's default: base 0
    free = S.build_dag("dialogue 4", DIALOGUE4, dialogue_index=3, freeze_dialogue_seed=None)
    assert free["C"][2].bg_seed == S.turn_seeds(3, 2)[0] == 3 * 4 + 2 + 3


def test_duplicates_follow_the_reference_rule():
    """theatergen.py:217-226: only a repeat of the IMMEDIATELY previous (prompt, obj_id) is skipped (its slot aliases
    the previous outputs); a non-consecutive repeat is generated again, by then with the id's reference image."""
    d = {"turn 1": {"objects": [["a fox", [0, 0, 100, 100], 1], ["a fox", [50, 0, 100, 100], 1],      # consecutive
                                ["an owl", [200, 0, 100, 100], 2], ["a fox", [300, 0, 100, 100], 1],  # non-consecutive
                                ["a fox", [300, 50, 100, 100], 5]]},                                   # same phrase, new id
         "turn 3": {"objects": [["a fox", [0, 0, 100, 100], 1]]}}                                      # turn 2 missing
    dag = S.build_dag("d", d, dialogue_index=2, freeze_dialogue_seed=None)
    assert [(i.turn, i.obj_index, i.obj_id) for i in dag["A"]] == [(0, 0, 1), (0, 2, 2), (0, 4, 5)]
    assert [(i.turn, i.obj_index, i.obj_id, i.ip_scale) for i in dag["B"]] == [(0, 3, 1, 0.4), (2, 0, 1, 0.4)]
    assert [(i.turn, i.obj_index, i.alias_of) for i in dag["alias"]] == [(0, 1, 0)]
    assert [i.phase for i in dag["turns"][0]] == ["A", "alias", "A", "B", "A"]       # every slot has a placeholder
    assert len(dag["C"]) == 2 and [c.turn for c in dag["C"]] == [0, 2]
    # `ind` counts turns that exist (generate.py:195-198 skips a missing turn before `ind += 1`): dialogue index 2 with
    # 4-turn predecessors -> turns_before 8, 9 ; ind_offset = seed_offset (0) + finished dialogues (2)
    assert [c.bg_seed for c in dag["C"]] == [8 + 2, 9 + 2]
    short = S.build_dag("d", d, dialogue_index=2, freeze_dialogue_seed=None, turns_before_dialogue=5)
    assert [c.bg_seed for c in short["C"]] == [5 + 2, 6 + 2]
    assert S.count_turns(d) == 2


def test_cmigbench_shapes_if_present():
    p = "/root/reference/CMIGBench/story.json"
    if not os.path.exists(p):
        pytest.skip("reference dataset not mounted (GPU box)")
    data = json.load(open(p))
    name = list(data)[3]
    dag = S.build_dag(name, data[name], 3)
    n_turns = sum(1 for t in range(4) if f"turn {t + 1}" in data[name])
    assert len(dag["C"]) == n_turns and len(dag["A"]) >= 1


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(0)
    ref = [torch.randn(s, generator=g) for s in ((5, 7), (3,), (2, 2, 2), (1000,))]
    mine = [t.clone() if rank == 0 else torch.zeros_like(t) for t in ref]
    S.broadcast_tensors(mine, src=0, bucket_bytes=64)     # tiny buckets: exercises the flush path
    ok = all(torch.equal(a, b) for a, b in zip(mine, ref))
    ids = [10, 11] if rank == 0 else [12]
    toks = torch.full((len(ids), 2, 4, 8), float(rank))
    allt = S.all_gather_ip_tokens(ids, toks)
    ok = ok and sorted(allt) == [10, 11, 12] and float(allt[12][0, 0, 0]) == 1.0 and float(allt[10][0, 0, 0]) == 0.0
    # weak-scaling bookkeeping used by bench.py: max over ranks of the per-rank time
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = ok and t.item() == float(world)
    ret[rank] = ok
    dist.destroy_process_group()


def test_collectives_gloo_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert all(ret[r] for r in range(world)), dict(ret)
