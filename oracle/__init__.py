"""oracle/ -- CPU restatement of TheaterGen's denoising hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may import
this package, and only as the checker / the timed CPU baseline.  Nothing under `theatergen_b200/`
imports it; the product path has no CPU fallback.

Parity status: the reference ships no numeric tests, golden vectors or fixtures for this path
("parity unpinned").  The pins that exist were generated in the build container by
importing the reference's own importable modules from /root/reference
(ip_adapter/attention_processor.py via a 6-line `diffusers.utils` shim, ip_adapter/resampler.py,
utils/guidance.py, utils/utils.py, utils/schedule.py) on seeded inputs -- see tests/golden/make_golden.py
and the committed tests/golden/*.pt.  The third-party arithmetic (diffusers 0.21.4 ResnetBlock2D,
Down/Upsample2D, Timesteps, TimestepEmbedding, DDIMScheduler) is restated from the published algorithm
and is anchored only on the reference's call sites; it stays "unpinned".

All functions are pure PyTorch, dtype/device agnostic (fp32 on CPU for parity, fp16 on a GPU to
reproduce the reference's own eager numerics), and cite the reference file:line they follow.
"""
