"""Host<->device copy rates from pinned memory and the NUMA placement of the GPU (explains the e2e gap of the
reduced-order configurations: 128 MB of parameter rows per 1e6 samples)."""
import glob, os, subprocess, time
import torch
print(subprocess.run("nvidia-smi topo -m; nproc; lscpu | grep -i 'numa\\|model name\\|socket'", shell=True, capture_output=True, text=True).stdout)
for p in glob.glob("/sys/bus/pci/devices/*/numa_node"):
    cls = open(os.path.join(os.path.dirname(p), "class")).read().strip()
    if cls.startswith("0x0302") or cls.startswith("0x0300"):
        print(p, open(p).read().strip(), open(os.path.join(os.path.dirname(p), "local_cpulist")).read().strip())
print("affinity", sorted(os.sched_getaffinity(0))[:8], "...", len(os.sched_getaffinity(0)))
dev = torch.device("cuda:0")
for mb in (1, 8, 128):
    n = mb * 1024 * 1024 // 8
    h = torch.empty(n, dtype=torch.float64).pin_memory()
    h.fill_(1.0)
    d = torch.empty(n, dtype=torch.float64, device=dev)
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name} {mb} MB pinned: {ms:.3f} ms  {mb / 1024 / (ms / 1e3):.1f} GB/s")
