"""CUDA-event timing decorator with the reference helper's interface
(flash_cosine_sim_attention/benchmark.py:7-58): `benchmark(fn, num_times=, warmup_iters=,
forwards=, backwards=)(*args)` -> mean milliseconds per call.  Backward is timed as
`fn(...).sum().backward()`, like the reference.  Used by the reference's top-level benchmark.py."""
import functools

import torch


def benchmark(fn, *, num_times=10, warmup_iters=10, forwards=True, backwards=False):
    assert forwards or backwards, "time at least one of forwards / backwards"

    @functools.wraps(fn)
    def timed(*args, **kwargs):
        for _ in range(warmup_iters):
            out = fn(*args, **kwargs)
            if backwards:
                out.sum().backward()
        total_ms = 0.0
        for _ in range(num_times):
            start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if forwards:
                start.record()
            out = fn(*args, **kwargs)
            if not forwards:
                start.record()          # backward only: clock starts after the forward
            if backwards:
                out.sum().backward()
            stop.record()
            torch.cuda.synchronize()
            total_ms += start.elapsed_time(stop)
        return total_ms / num_times

    return timed
