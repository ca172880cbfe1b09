"""torchrun --nproc-per-node N scripts/switch_probe.py: NVSwitch multicast all-reduce (csrc/allreduce_mc.cu) against NCCL
on the gradient arena's sizes: correctness, then time alone (CUDA events, max over ranks) per CTA count."""
import datetime, os, sys
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from audio8_b200.parallel import SwitchAllReduce
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
world = dist.get_world_size()
n = 90448384 + 4596224  # wav2vec2-base arena: early region + tail (fp32 elements)
import torch.distributed._symmetric_memory as symm
if rank == 0:
    print("multicast support:", symm._SymmetricMemory.has_multicast_support(torch._C._autograd.DeviceType.CUDA, local), flush=True)
sw = SwitchAllReduce(n, dev, None)
if rank == 0:
    print(f"world {world}, multicast ptr {sw.mc:#x}", flush=True)
g = torch.Generator(device=dev).manual_seed(rank)
src = torch.randn(n, device=dev, generator=g)
ref = src.clone()
dist.all_reduce(ref, op=dist.ReduceOp.AVG)
sw.buf.copy_(src)
torch.cuda.synchronize(); dist.barrier()
sw.start(0, 1000000).wait(); sw.start(1000000, n).wait()
torch.cuda.synchronize()
err = (sw.buf - ref).abs().max().item()
same = [torch.empty(4096, device=dev) for _ in range(world)]
dist.all_gather(same, sw.buf[12345:12345 + 4096].contiguous())
print(f"rank {rank}: max |switch - nccl| = {err:.3g} (scale {ref.abs().max().item():.3g}); identical on all ranks: "
      f"{all(torch.equal(s, same[0]) for s in same)}", flush=True)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


for lo, hi, name in ((0, 90448384, "early region 362 MB"), (90448384, n, "tail 18 MB")):
    t = timed(lambda: dist.all_reduce(ref[lo:hi], op=dist.ReduceOp.AVG))
    if rank == 0:
        print(f"{name}: NCCL {t * 1e3:.0f} us ({(hi - lo) * 4 / t / 1e6:.0f} GB/s algorithmic)", flush=True)
    for ctas in (8, 16, 24, 32, 48, 64):
        sw.ctas = ctas
        t = timed(lambda: sw.start(lo, hi).wait())
        if rank == 0:
            print(f"{name}: switch, {ctas} CTAs {t * 1e3:.0f} us ({(hi - lo) * 4 / t / 1e6:.0f} GB/s algorithmic)", flush=True)
    t = timed(lambda: torch.ops.symm_mem.multimem_all_reduce_(sw.buf[lo:hi], "sum", dist.group.WORLD.group_name))
    if rank == 0:
        print(f"{name}: torch multimem_all_reduce_ {t * 1e3:.0f} us", flush=True)
dist.destroy_process_group()
