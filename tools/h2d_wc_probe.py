"""Does write-combined pinned host memory move the host-copy ceiling?  Plain cudaMemcpyAsync H2D (+ D2H into
default pinned memory) of one bench step's bytes, input buffer allocated with cudaHostAllocDefault vs
cudaHostAllocWriteCombined.  `python tools/h2d_wc_probe.py` (one GPU; run under torchrun for N ranks)."""
import ctypes
import os
import time

import torch

rank = int(os.environ.get("LOCAL_RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
rt = ctypes.CDLL("libcudart.so.12")
IN_BYTES, OUT_BYTES = 256 * 524160 * 4, 256 * 4096 * 80 * 4
d_in = torch.empty(IN_BYTES, dtype=torch.uint8, device="cuda")
d_out = torch.empty(OUT_BYTES, dtype=torch.uint8, device="cuda")
h_out = torch.empty(OUT_BYTES, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(device_ids=[rank])


for name, flag in (("default", 0), ("write_combined", 4)):
    p = ctypes.c_void_p()
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(IN_BYTES), ctypes.c_uint(flag)) == 0
    ctypes.memset(p, 1, IN_BYTES)
    for both in (False, True):
        def step():
            rt.cudaMemcpyAsync(ctypes.c_void_p(d_in.data_ptr()), p, ctypes.c_size_t(IN_BYTES), 1, ctypes.c_void_p(s1.cuda_stream))
            if both:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        if world > 1:
            t = torch.tensor([dt], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t)
        if rank == 0:
            print(f"{world} rank(s), input {name:15s} {'H2D + D2H' if both else 'H2D only '}: {dt * 1e3:7.2f} ms per step, "
                  f"H2D {IN_BYTES / dt / 1e9:5.1f} GB/s per rank" + (f", D2H {OUT_BYTES / dt / 1e9:5.1f} GB/s per rank" if both else ""), flush=True)
    rt.cudaFreeHost(p)
