"""Where does pinned host memory live relative to the GPU, and what does that cost on the PCIe link?

Prints the GPU's NUMA node (sysfs), the process's CPU / memory-node masks, and the pinned H2D / D2H bandwidth of a
21.5 MB transfer (the configs[1] step input) for pinned buffers allocated under set_mempolicy(MPOL_PREFERRED, node) for
every memory node.  Diagnostic for the end-to-end leg of bench.py (box-to-box spread 3.7 .. 8.2 M solves/s)."""
import ctypes
import glob
import json
import os
import sys

import numpy as np
import torch

libc = ctypes.CDLL(None, use_errno=True)
SYS_set_mempolicy, SYS_get_mempolicy, SYS_move_pages = 238, 239, 279   # x86_64
MPOL_DEFAULT, MPOL_PREFERRED, MPOL_BIND = 0, 1, 2


def set_mempolicy(mode, node=None):
    if node is None:
        r = libc.syscall(SYS_set_mempolicy, MPOL_DEFAULT, None, 0)
    else:
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        r = libc.syscall(SYS_set_mempolicy, mode, mask, 1024)
    return r, (ctypes.get_errno() if r != 0 else 0)


def page_nodes(t, n=8):
    """NUMA node of the first n pages of a (pinned) tensor via move_pages(nodes=NULL)."""
    ps = 4096
    base = t.data_ptr() & ~(ps - 1)
    pages = (ctypes.c_void_p * n)(*[base + i * ps * max(1, t.numel() * t.element_size() // ps // n) for i in range(n)])
    status = (ctypes.c_int * n)()
    r = libc.syscall(SYS_move_pages, 0, n, pages, None, status, 0)
    return list(status) if r == 0 else ("move_pages errno %d" % ctypes.get_errno())


def bw(h, d, reps=30):
    s = torch.cuda.current_stream()
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.synchronize(); e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record(); s.synchronize()
    up = h.numel() * h.element_size() * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    s.synchronize(); e0.record()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    e1.record(); s.synchronize()
    dn = h.numel() * h.element_size() * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    return up, dn


def main():
    out = {}
    torch.cuda.init()
    prop = torch.cuda.get_device_properties(0)
    bus = "%04x:%02x:%02x.0" % (prop.pci_domain_id, prop.pci_bus_id, prop.pci_device_id) if hasattr(prop, "pci_bus_id") else None
    if bus is None:
        import subprocess
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", "0"], capture_output=True,
                             text=True).stdout.strip().lower()[4:]
    out["gpu_bus"] = bus
    try:
        out["gpu_numa_node"] = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read())
    except Exception as e:
        out["gpu_numa_node"] = "unreadable: %r" % (e,)
    nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
    out["nodes"] = {n: open("/sys/devices/system/node/node%d/cpulist" % n).read().strip() for n in nodes}
    out["affinity"] = "%d cpus: %s..." % (len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8])
    for k in ("Cpus_allowed_list", "Mems_allowed_list"):
        for line in open("/proc/self/status"):
            if line.startswith(k):
                out[k] = line.split(":", 1)[1].strip()
    out["cpu_now"] = libc.sched_getcpu()
    nbytes = 21495808
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    res = {}
    h = torch.empty(nbytes + 4096, dtype=torch.uint8).pin_memory()
    res["default"] = dict(pages=page_nodes(h), bw=bw(h[:nbytes], d))
    for n in nodes:
        r = set_mempolicy(MPOL_PREFERRED, n)
        h = torch.empty(nbytes + 8192 * (n + 2), dtype=torch.uint8).pin_memory()
        set_mempolicy(MPOL_DEFAULT)
        res["preferred_node%d" % n] = dict(set_mempolicy=r, pages=page_nodes(h), bw=bw(h[:nbytes], d))
    # affinity route: run on a CPU of node n while allocating
    full = os.sched_getaffinity(0)
    for n in nodes:
        cpus = set()
        for part in out["nodes"][n].split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            cpus |= set(range(int(a), int(b or a) + 1))
        cpus &= full
        if not cpus:
            res["affinity_node%d" % n] = "no allowed cpu on this node"
            continue
        os.sched_setaffinity(0, cpus)
        h = torch.empty(nbytes + 8192 * (n + 40), dtype=torch.uint8).pin_memory()
        res["affinity_node%d" % n] = dict(cpu=libc.sched_getcpu(), pages=page_nodes(h), bw=bw(h[:nbytes], d))
        os.sched_setaffinity(0, full)
    # the library's allocator (mmap + mbind to the GPU's node + cudaHostRegister), default and forced to every node
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from racinglmpc_b200 import _native as nat
    for tag, env in [("lmpc_host_alloc_auto", None)] + [("lmpc_host_alloc_node%d" % n, str(n)) for n in nodes] + [("lmpc_host_alloc_off", "off")]:
        if env is None:
            os.environ.pop("LMPC_B200_NUMA", None)
        else:
            os.environ["LMPC_B200_NUMA"] = env
        a = nat.pinned_empty(nbytes, np.uint8, device=0)
        h = torch.from_numpy(a)
        res[tag] = dict(node=int(nat.lib().lmpc_host_numa_node(0)), pages=nat.page_nodes(a), bw=bw(h, d))
    os.environ.pop("LMPC_B200_NUMA", None)
    out["h2d_d2h_gbs"] = res
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
