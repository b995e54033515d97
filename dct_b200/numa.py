"""NUMA placement for the host side of the path (end-to-end multi-GPU throughput is
bound by PCIe and host DRAM, so the pinned staging buffers of a rank must live on the socket its GPU
hangs off).  Linux places pages on the node of the CPU that first touches them, so binding the
process to the GPU-local cores BEFORE allocating pinned memory is enough.  Pure host plumbing."""
from __future__ import annotations

import os


def _cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(pci_bus_id: str) -> int | None:
    """NUMA node of a GPU given its PCI bus id ('00000000:1B:00.0' or '0000:1b:00.0')."""
    dom, bus, rest = pci_bus_id.strip().lower().split(":")
    path = f"/sys/bus/pci/devices/{dom[-4:]}:{bus}:{rest}/numa_node"
    try:
        node = int(open(path).read())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def bind_to_gpu_node(device_index: int) -> dict:
    """Restrict this process to the cores of the NUMA node of CUDA device `device_index`.
    Returns what was done ({} if nothing could be determined); never raises."""
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device_index).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = gpu_numa_node(bus)
        if node is None:
            return {}
        cpus = _cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return {}
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed), "pci": bus}
    except Exception:  # plumbing only: a box without sysfs / NVML simply stays unbound
        return {}
