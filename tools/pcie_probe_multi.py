"""Host <-> device ceiling of ONE box with N ranks working at the same time (the question behind bench.py's `e2e` at N > 1):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/pcie_probe_multi.py

Every rank, concurrently (barrier before each measurement, device time by CUDA events, max over ranks):
  * cudaMemcpyAsync from / to pinned memory, 4 MB and 64 MB, each direction and both at once;
  * the approxNDCG host entry (`ltr_approx_ndcg_fwdbwd_host`, 4096 x 128 per step, 3 streams in flight) with the kernel
    reading pinned host memory itself (zero copy) and with staged copies (LTR_HOST_ZERO_COPY=0 is read once per process,
    so the staged arm runs in a child interpreter).
Rank 0 prints one JSON line: per-rank and aggregate GB/s, and what the e2e path reaches against them."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200")]
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, reps):
        fn(); fn()
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / reps * 1e-3], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    out = {"n_gpus": world}
    if os.environ.get("LTR_PROBE_ARM") != "staged":
        for mb in (4, 64):
            n = mb * 1024 * 1024 // 4
            h, h2 = torch.empty(n).pin_memory(), torch.empty(n).pin_memory()
            d, d2 = torch.empty(n, device=dev), torch.empty(n, device=dev)
            s2 = torch.cuda.Stream()
            reps = 200 if mb == 4 else 30
            t = timed(lambda: d.copy_(h, non_blocking=True), reps)
            out[f"h2d_{mb}MB_GBs_per_rank"] = mb / 1024 / t
            t = timed(lambda: h2.copy_(d2, non_blocking=True), reps)
            out[f"d2h_{mb}MB_GBs_per_rank"] = mb / 1024 / t

            def both():
                d.copy_(h, non_blocking=True)
                s2.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s2):
                    h2.copy_(d2, non_blocking=True)
                torch.cuda.current_stream().wait_stream(s2)
            t = timed(both, reps)
            out[f"both_{mb}MB_GBs_per_rank"] = 2 * mb / 1024 / t
    # the e2e path itself
    from ltr_b200.host import ApproxNDCGHost
    B, n = 4096, 128
    g = torch.Generator().manual_seed(rank)
    ys = [torch.randint(0, 5, (B, n), generator=g).float().pin_memory() for _ in range(6)]
    ss = [torch.randn(B, n, generator=g).pin_memory() for _ in range(6)]
    plan = ApproxNDCGHost(B, n, alpha=10.0, device=dev, depth=3)
    state = {"i": 0}

    def step():
        i = state["i"]; state["i"] = i + 1
        plan.submit(ss[i % 6], ys[i % 6])
    for _ in range(6):
        step()
    plan.wait_all()
    sync_all()
    import time
    t0 = time.perf_counter()
    steps = 300
    for _ in range(steps):
        step()
    plan.wait_all()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    arm = "staged" if os.environ.get("LTR_HOST_ZERO_COPY") == "0" else "zero_copy"
    out[f"e2e_{arm}_slates_per_s_total"] = world * B * steps / float(dt)
    out[f"e2e_{arm}_GBs_per_rank_h2d_plus_d2h"] = (3 * B * n * 4 + 4) * steps / float(dt) / 1e9
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()


if __name__ == "__main__":
    main()
