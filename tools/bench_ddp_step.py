"""BASELINE configs[1] as a data-parallel TRAINING STEP on N GPUs of one box (the reference's DDP loop,
main_batch_execution.py:120-171): every rank takes its own 4096 slates x 128 docs x 136 features.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 tools/bench_ddp_step.py

  fused  : `ltr_triple_approx_ndcg_step` (fold, TMA + tcgen05 kernel, reduce, unfold) writes the six gradients into ONE flat
           buffer, then ONE NCCL all-reduce of that buffer (13.4 k floats) averages them over the ranks -- what DDP does.
  fused_p2p : the same step followed by `ltr_peer_allreduce_avg` (two small launches: push into every peer's inbox over NVLink,
           gather in rank order) instead of NCCL.
  stock  : DistributedDataParallel(TripleLayerNet) + the K1 approxNDCG loss kernel + loss.backward() (fp32 and tf32 matmuls).
CUDA events over 30 steps after warm-up, barrier on both sides, max over ranks; rank 0 prints one JSON line with whole-job
slates/s of each arm and the check that the averaged fused gradients equal the DDP ones."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "nn-with-pytorch-personalized-losses_b200"), os.path.join(ROOT, "tests")]
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from ltr_b200 import _lib, _runtime as rt
    from ltr_b200.architeture.tripleLayer import TripleLayerNet
    from ltr_b200.losses.approxNDCG import approxNDCGLoss
    from ltr_testing import make_slates
    lib = _lib.load()
    B, n, F, ring, iters = 4096, 128, 136, 3, 30
    torch.manual_seed(0)                                             # the same initial weights on every rank
    net = TripleLayerNet(F).to(dev)
    data = []
    for i in range(ring):
        g = torch.Generator().manual_seed(1000 * rank + i)
        data.append((torch.rand(B, n, F, generator=g).to(dev), make_slates(B, n, seed=1000 * rank + 50 + i, min_len=64)[1].to(dev)))
    prm = [t.detach() for t in (net.l1.weight, net.l1.bias, net.l2.weight, net.l2.bias, net.l3.weight, net.l3.bias)]
    sizes = [t.numel() for t in prm]
    flat = torch.zeros(sum(sizes), device=dev)
    offs, o = [], 0
    for s in sizes:
        offs.append(o); o += s
    scores, slate, loss = torch.empty(B, n, device=dev), torch.empty(B, device=dev), torch.empty(1, device=dev)
    nscr = lib.ltr_triple_step_scratch_bytes(B, n, F)
    scr = torch.empty(nscr, dtype=torch.uint8, device=dev)
    tab, ws = rt.disc_prefix(dev), rt.workspace(dev)
    base = flat.data_ptr()

    def fused(i):
        X, y = data[i % ring]
        rc = lib.ltr_triple_approx_ndcg_step(X.data_ptr(), y.data_ptr(), *[t.data_ptr() for t in prm], B, n, F, 10.0, 1e-10, -1.0,
                                             tab.data_ptr(), tab.numel(), scores.data_ptr(), slate.data_ptr(), loss.data_ptr(),
                                             *[base + 4 * o for o in offs], scr.data_ptr(), nscr, ws.data_ptr(), rt.stream_ptr())
        assert rc == 0
        if world > 1:
            dist.all_reduce(flat)                                    # SUM over ranks; the 1 / world of DDP's average is applied below
        flat.mul_(1.0 / world)

    from ltr_b200 import sharding
    ex = sharding.grad_peer_exchange(dist.group.WORLD, dev) if world > 1 else None

    def fused_p2p(i):
        X, y = data[i % ring]
        rc = lib.ltr_triple_approx_ndcg_step(X.data_ptr(), y.data_ptr(), *[t.data_ptr() for t in prm], B, n, F, 10.0, 1e-10, -1.0,
                                             tab.data_ptr(), tab.numel(), scores.data_ptr(), slate.data_ptr(), loss.data_ptr(),
                                             *[base + 4 * o for o in offs], scr.data_ptr(), nscr, ws.data_ptr(), rt.stream_ptr())
        assert rc == 0
        if ex is not None:
            ex.average([flat])

    ddp = DDP(net, device_ids=[local]) if world > 1 else net

    def stock(i):
        X, y = data[i % ring]
        net.zero_grad(set_to_none=True)
        approxNDCGLoss(ddp(X, None, None).squeeze(-1), y, alpha=10.0).backward()

    def timed(fn, iters):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(iters):
            fn(5 + i)
        b.record()
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([a.elapsed_time(b) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    ms_fused = timed(fused, iters)
    ms_p2p = timed(fused_p2p, iters)
    us_nccl_alone = 1e3 * timed(lambda i: dist.all_reduce(flat), 100) if world > 1 else 0.0
    us_p2p_alone = 1e3 * timed(lambda i: ex.average([flat]), 100) if ex is not None else 0.0

    def graphed(fn, per_graph=20, replays=10):
        """device time per call with the host out of the picture: `per_graph` calls captured in a CUDA graph"""
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                fn(0)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    for i in range(per_graph):
                        fn(i)
            torch.cuda.current_stream().wait_stream(side)
            return 1e3 * timed(lambda i: g.replay(), replays) / per_graph
        except Exception as e:                                        # (reported, not fatal: this is a side measurement)
            return "failed: %s" % str(e).splitlines()[0][:80]

    us_p2p_graph = graphed(lambda i: ex.average([flat])) if ex is not None else 0.0
    us_step_p2p_graph = graphed(fused_p2p) if ex is not None else 0.0
    fused(0)
    nccl_avg = flat.clone()
    fused_p2p(0)
    p2p_err = float((flat - nccl_avg).abs().max() / nccl_avg.abs().max())
    p2p_status = ex.status() if ex is not None else 0
    torch.backends.cuda.matmul.allow_tf32 = False
    ms_fp32 = timed(stock, 10)
    torch.backends.cuda.matmul.allow_tf32 = True
    ms_tf32 = timed(stock, 10)
    torch.backends.cuda.matmul.allow_tf32 = False
    # parity of the averaged gradients on one common step
    fused(0)
    mine = [flat[o:o + s].clone() for o, s in zip(offs, sizes)]
    stock(0)
    err = max(float((m - p.grad.reshape(-1)).abs().max() / p.grad.abs().max().clamp_min(1e-30))
              for m, p, name in zip(mine, net.parameters(), "abcdef") if name != "f")      # (l3.bias: ~0, see the tests)
    if rank == 0:
        print(json.dumps({"metric": "TripleLayerNet + approxNDCG data-parallel training step, slates/s (n=128, F=136)", "n_gpus": world,
                          "slates_per_rank": B, "fused_ms_per_step": ms_fused, "fused_slates_per_s": world * B / (ms_fused * 1e-3),
                          "fused_p2p_ms_per_step": ms_p2p, "fused_p2p_slates_per_s": world * B / (ms_p2p * 1e-3), "p2p_used": ex is not None,
                          "p2p_status": p2p_status,
                          "allreduce_alone_us": {"nccl": us_nccl_alone, "peer_memory": us_p2p_alone, "peer_memory_in_cuda_graph": us_p2p_graph},
                          "fused_p2p_step_in_cuda_graph_us": us_step_p2p_graph, "p2p_vs_nccl_max_rel_diff": p2p_err,
                          "ddp_torch_fp32_ms_per_step": ms_fp32, "ddp_torch_fp32_slates_per_s": world * B / (ms_fp32 * 1e-3),
                          "ddp_torch_tf32_ms_per_step": ms_tf32, "ddp_torch_tf32_slates_per_s": world * B / (ms_tf32 * 1e-3),
                          "speedup_vs_ddp_fp32": ms_fp32 / ms_fused, "speedup_vs_ddp_tf32": ms_tf32 / ms_fused,
                          "speedup_p2p_vs_ddp_fp32": ms_fp32 / ms_p2p, "speedup_p2p_vs_ddp_tf32": ms_tf32 / ms_p2p,
                          "max_rel_err_of_averaged_gradients_vs_ddp_fp32": err,
                          "collective": "fused: one NCCL all-reduce of 13,441 floats per step; fused_p2p: ltr_peer_allreduce_avg (one launch, NVLink peer stores); stock: DDP's bucket all-reduce"}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
