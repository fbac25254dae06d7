"""Gradient all-reduce alternatives at the trainer's size (6.15 M fp32 = 24.6 MB), one process per GPU:
NCCL all_reduce vs torch symmetric-memory two-shot / multimem (NVLS) all-reduce over NVLink peer memory.
Correctness against NCCL, time per call (CUDA events, max over ranks), eager and replayed from a CUDA graph.
usage: torchrun --nproc-per-node N tools/allreduce_ab.py [numel]"""
import os, sys
import torch
import torch.distributed as dist

n = int(sys.argv[1]) if len(sys.argv) > 1 else 6_150_000
rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank)))
dist.init_process_group('nccl')
dev = torch.device('cuda')
pad = 128 * world
npad = (n + pad - 1) // pad * pad
g = torch.Generator(device='cuda').manual_seed(100 + rank)
src = torch.randn(npad, device=dev, generator=g)
ref = src.clone()
dist.all_reduce(ref)


def timed(fn, iters=30):
    for _ in range(5):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters * 1e3], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


buf = src.clone()
res = {'nccl': timed(lambda: dist.all_reduce(buf))}
variants = {}
try:
    import torch.distributed._symmetric_memory as symm_mem
    gname = dist.group.WORLD.group_name
    sbuf = symm_mem.empty(npad, dtype=torch.float32, device=dev)
    hdl = symm_mem.rendezvous(sbuf, dist.group.WORLD)
    variants['two_shot'] = lambda: torch.ops.symm_mem.two_shot_all_reduce_(sbuf, 'sum', gname)
    if getattr(hdl, 'multicast_ptr', 0):
        variants['multimem'] = lambda: torch.ops.symm_mem.multimem_all_reduce_(sbuf, 'sum', gname)
    else:
        res['multimem'] = 'no multicast support'
except Exception as e:                                   # noqa: BLE001 -- report, this is a probe
    res['symm_mem'] = f'unavailable: {type(e).__name__}: {e}'
for name, fn in variants.items():
    try:
        sbuf.copy_(src)
        dist.barrier(); torch.cuda.synchronize()
        fn()
        torch.cuda.synchronize()
        err = ((sbuf - ref).abs().max() / ref.abs().max()).item()
        same = sbuf.clone()
        dist.broadcast(same, src=0)
        identical = bool(torch.equal(same, sbuf))          # every rank holds the same bits as rank 0
        t = timed(fn)
        # replayed from a CUDA graph (what the trainer does)
        sbuf.copy_(src)
        dist.barrier(); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            fn()
        sbuf.copy_(src)
        dist.barrier(); torch.cuda.synchronize()
        gr.replay(); torch.cuda.synchronize()
        gerr = ((sbuf - ref).abs().max() / ref.abs().max()).item()
        tg = timed(gr.replay)
        res[name] = dict(us=round(t, 1), graph_us=round(tg, 1), rel_err=err, graph_rel_err=gerr, identical_across_ranks=identical)
    except Exception as e:                               # noqa: BLE001
        res[name] = f'failed: {type(e).__name__}: {e}'
if rank == 0:
    print(f'ALLREDUCE world {world} numel {npad} ({npad * 4 / 1e6:.1f} MB): {res}', flush=True)
dist.barrier()
os._exit(0)
