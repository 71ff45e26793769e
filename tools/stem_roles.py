"""Where the conv-stem kernel's roles wait (needs a -DBHSTEM_PROFILE build of libbhstem.so):

    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared \
         -DBHSTEM_PROFILE -o build/libbhstem_prof.so beatheritage_b200/csrc/bhstem.cu
    BHSTEM_LIB=$PWD/build/libbhstem_prof.so python tools/stem_roles.py [B]

Prints, per stage, the share of the kernel's cycles each role spent blocked on each barrier."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from beatheritage_b200 import _stem_lib  # noqa: E402
from beatheritage_b200.conv_stem import ConvStem  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
torch.manual_seed(0)
stem = ConvStem(464, 768).to(dev)
x = (torch.randn(B, 4096, 464, device=dev) * 1.5).to(torch.bfloat16)
h = stem.forward_stage(1, x)
lib = _stem_lib.lib()
buf = (ctypes.c_ulonglong * 12)()
names = ["producer on A-empty", "producer on W-empty", "MMA on A-full", "MMA on W-full", "MMA on TMEM-empty",
         "epilogue on TMEM-full"]
for stage, inp in ((1, x), (2, h)):
    for _ in range(3):
        stem.forward_stage(stage, inp)
    torch.cuda.synchronize()
    lib.bhstem_debug_profile(buf)
    stem.forward_stage(stage, inp)
    torch.cuda.synchronize()
    lib.bhstem_debug_profile(buf)
    ctas, total = buf[7], buf[6]
    print(f"stage {stage}: {ctas} CTAs, {total / ctas:.0f} cycles per CTA, SM clock during the kernel "
          f"{1e3 * buf[8] / max(buf[9], 1):.0f} MHz (clock64 / globaltimer of CTA 0)")
    for i, n in enumerate(names):
        print(f"   {n:24s} {100.0 * buf[i] / total:5.1f} %")

# clocks and power while the stem runs back to back (is the tensor pipe power-capped?)
try:
    import pynvml
    pynvml.nvmlInit()
    hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
    for _ in range(400):
        stem(x)
    samples = []
    for _ in range(20):
        samples.append((pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM),
                        pynvml.nvmlDeviceGetPowerUsage(hnd) / 1000.0))
    torch.cuda.synchronize()
    samples.sort()
    print(f"under load: SM clock median {samples[len(samples) // 2][0]} MHz, power max {max(p for _, p in samples):.0f} W")
except Exception as e:  # noqa: BLE001
    print("nvml unavailable:", e)
