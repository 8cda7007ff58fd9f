"""Dev experiment (2+ GPUs, torchrun): bandwidth of a rank pushing into the root's IPC-mapped buffer with a raw peer
cudaMemcpyAsync vs torch's cross-device copy_ vs an NCCL send/recv, and whether the push disturbs a kernel on the root."""
import ctypes, os, sys, time
import torch, torch.distributed as dist
from torch.multiprocessing.reductions import reduce_tensor

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = "cuda:%d" % local
dist.init_process_group("nccl", device_id=torch.device(dev))
n = 1 << 27  # doubles = 1 GiB
src = torch.ones(n, dtype=torch.float64, device=dev)
recv = torch.zeros((world, n), dtype=torch.float64, device=dev) if rank == 0 else None
obj = [reduce_tensor(recv) if rank == 0 else None]
dist.broadcast_object_list(obj, src=0)
peer = None
if rank != 0:
    fn, a = obj[0]
    peer = fn(*a)
    peer[rank, :1].copy_(src[:1]); torch.cuda.synchronize()
    print("rank", rank, "can access peer 0:", torch.cuda.can_device_access_peer(local, 0), flush=True)
cudart = ctypes.CDLL("libcudart.so.12")
cudart.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
st = torch.cuda.Stream(device=dev)
dist.barrier(); torch.cuda.synchronize()
def timeit(f, reps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f(); torch.cuda.synchronize()
    e0.record(st)
    for _ in range(reps): f()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
if rank == 1:
    ms = timeit(lambda: cudart.cudaMemcpyAsync(peer[rank].data_ptr(), src.data_ptr(), n * 8, 4, st.cuda_stream))
    print("raw peer cudaMemcpyAsync: %.2f ms per GiB = %.1f GB/s" % (ms, n * 8 / ms / 1e6), flush=True)
    def tc():
        with torch.cuda.stream(st):
            peer[rank].copy_(src, non_blocking=True)
    ms = timeit(tc)
    print("torch copy_: %.2f ms per GiB = %.1f GB/s" % (ms, n * 8 / ms / 1e6), flush=True)
dist.barrier(); torch.cuda.synchronize()
# NCCL send / recv of the same GiB
for rep in range(3):
    torch.cuda.synchronize(); dist.barrier(); t = time.time()
    if rank == 1: dist.send(src, dst=0)
    if rank == 0: dist.recv(recv[1], src=1)
    torch.cuda.synchronize(); dt = time.time() - t
    if rank == 0: print("nccl send/recv: %.2f ms per GiB = %.1f GB/s" % (dt * 1e3, n * 8 / dt / 1e9), flush=True)
# does a push slow a compute kernel on the root?  root runs a long elementwise chain, rank 1 pushes meanwhile
x = torch.rand(1 << 26, dtype=torch.float64, device=dev)
def work():
    y = x
    for _ in range(40): y = y * 1.0000001 + 1e-9
    return y
for mode in ("idle", "push"):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if rank == 0: work()
    if rank == 1 and mode == "push":
        for _ in range(4): cudart.cudaMemcpyAsync(peer[rank].data_ptr(), src.data_ptr(), n * 8, 4, st.cuda_stream)
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print("root compute while peer is %s: %.2f ms" % (mode, e0.elapsed_time(e1)), flush=True)
dist.barrier()
