"""BASELINE.json configs[3] / SURVEY.md 8e: the CMIGBench "dialogue 4" DAG (4 turns, 10 character generations, 4 ids)
executed by theatergen_b200.runner.DagRunner -- phases A / B / C with the IP-token and stage-1 exchanges -- on ONE rank
and on TWO ranks: every stage-1 trajectory and every stage-2 result must be bit-identical whatever the world size
(seeds are a function of (dialogue, turn) only, co-batches have a fixed size, kernels are batch-position independent).

Two ranks: NCCL on two GPUs when the box has them (gpurun --gpus 2); otherwise both ranks share GPU 0 and exchange
through gloo (CPU staging) -- the same runner code apart from the staging copy."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
STEPS = 4


def _build(device):
    from oracle import controlnet as OC
    from oracle import synth
    from oracle import unet as OU
    from theatergen_b200.controlnet import TGControlNet
    from theatergen_b200.engine import DenoiseEngine
    from theatergen_b200.ip_adapter import IPAdapter
    from theatergen_b200.ip_adapter.ip_adapter import _Pipe
    from theatergen_b200.runner import DagRunner
    from theatergen_b200.scheduler import DDIMScheduler
    from theatergen_b200.unet import TGUNet
    net = TGUNet(OU.synth_state_dict(1234), device)
    cn = TGControlNet(OC.synth_state_dict(4321), device)
    sched = DDIMScheduler()
    adapter = IPAdapter(_Pipe(net, sched), None, None, device)
    adapter.image_proj_model.load_state_dict(synth.fill(synth.image_proj_shapes(), 9), device=device)
    eng = DenoiseEngine(net, sched, controlnet=cn)
    return DagRunner(eng, adapter, cobatch=2, num_steps=STEPS)


def _dialogue():
    d = json.load(open(os.path.join(HERE, "golden", "cmigbench_dialogue4.json")))
    return "dialogue 4", d["dialogue 4"]


def _run(runner):
    name, dlg = _dialogue()
    s1, s2, stats = runner.run_dialogue(name, dlg, dialogue_index=3)
    return ({k: v.cpu() for k, v in s1.items()}, {k: (v[0].cpu(), v[1].cpu()) for k, v in s2.items()}, stats)


def _worker(rank, world, port, backend, out_dir):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    if backend == "nccl":
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    s1, s2, stats = _run(_build(dev))
    torch.save(dict(s1=s1, s2=s2, steps=stats.denoise_steps, padded=stats.padded_steps), os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_dialogue4_dag_one_rank_vs_two_ranks_bit_exact(tmp_path):
    import torch.multiprocessing as mp
    from theatergen_b200 import sharding
    name, dlg = _dialogue()
    dag = sharding.build_dag(name, dlg, 3)
    assert [len(dag[p]) for p in "ABC"] == [4, 6, 4] and not dag["alias"]
    s1, s2, stats = _run(_build(torch.device("cuda", 0)))
    assert len(s1) == 10 and sorted(s2) == [0, 1, 2, 3]
    assert stats.denoise_steps == {"A": 4 * STEPS, "B": 6 * STEPS, "C": 4 * STEPS} and sum(stats.padded_steps.values()) == 0
    for v in s1.values():
        assert v.shape == (1, 4, 64, 64) and torch.isfinite(v.float()).all()
    # later appearances use the id's image tokens at scale 0.4: not the first appearance's result
    assert not torch.equal(s1[(0, 1)], s1[(2, 1)])
    # stage 2 keeps the frozen region: inside the composed layout mask the result is the re-noised composite
    lat, fg = s2[3]
    assert (fg != 0).any() and (fg == 0).any()

    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, backend, str(tmp_path)), nprocs=2, join=True)
    got = [torch.load(tmp_path / f"rank{r}.pt", weights_only=False) for r in range(2)]
    assert sum(sum(g["steps"].values()) for g in got) == 14 * STEPS       # every item ran exactly once
    for r in range(2):                                                    # every rank holds every stage-1 result
        assert sorted(got[r]["s1"]) == sorted(s1)
        for k in s1:
            assert torch.equal(got[r]["s1"][k], s1[k]), (backend, r, k)
    turns = {}
    for g in got:
        turns.update(g["s2"])
    assert sorted(turns) == [0, 1, 2, 3]
    for t in turns:
        assert torch.equal(turns[t][0], s2[t][0]) and torch.equal(turns[t][1], s2[t][1]), (backend, t)
    print(f"dialogue-4 DAG: 1 rank == 2 ranks ({backend}) bit for bit; padded steps on 2 ranks:",
          [g["padded"] for g in got])
