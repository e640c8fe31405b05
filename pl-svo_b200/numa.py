"""Host-side placement: run the calling process on the CPUs next to a GPU, so that the page-locked buffers it allocates
afterwards (first touch) and the threads that fill them live on the GPU's own NUMA node.

Why it matters here: the end-to-end path is bound by the host->device copy of the frame pairs,
and on a two-socket host a pinned buffer that sits on the far socket is copied at well under half the rate of a local one
(`tools/h2d_peak.py`).  Nothing in the reference corresponds to this; it is plumbing around the C ABI, used by `bench.py`
(every rank binds to its own GPU's node before allocating) and available to any integrating process.

Linux only; every function degrades to a no-op (returning None) when the topology cannot be read."""
from __future__ import annotations

import os


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def device_pci_bus_id(device_index: int) -> str | None:
    """'0000:1b:00.0' of CUDA device `device_index` (as torch numbers it in this process)."""
    try:
        import torch

        p = torch.cuda.get_device_properties(device_index)
        return f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        return None


def device_numa_node(device_index: int) -> int | None:
    bus = device_pci_bus_id(device_index)
    if not bus:
        return None
    try:
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def device_cpus(device_index: int) -> set[int] | None:
    """CPUs local to the GPU: sysfs `local_cpulist` of its PCI function, else the cpulist of its NUMA node, else NVML."""
    bus = device_pci_bus_id(device_index)
    if bus:
        for path in (f"/sys/bus/pci/devices/{bus}/local_cpulist",):
            try:
                cpus = _parse_cpulist(open(path).read())
                if cpus:
                    return cpus
            except Exception:
                pass
    node = device_numa_node(device_index)
    if node is not None:
        try:
            return _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read()) or None
        except Exception:
            pass
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode()) if bus else pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        return cpus or None
    except Exception:
        return None


def bind_to_device(device_index: int):
    """Restrict the calling process to the CPUs local to the GPU.  Returns the previous affinity mask (pass it to
    `restore`) or None if nothing was changed."""
    if not hasattr(os, "sched_setaffinity"):
        return None
    cpus = device_cpus(device_index)
    if not cpus:
        return None
    try:
        old = os.sched_getaffinity(0)
        allowed = cpus & old
        if not allowed or allowed == old:
            return None
        os.sched_setaffinity(0, allowed)
        return old
    except Exception:
        return None


def restore(mask) -> None:
    if mask:
        try:
            os.sched_setaffinity(0, mask)
        except Exception:
            pass


def describe(device_index: int) -> dict:
    cpus = device_cpus(device_index)
    return {"pci_bus_id": device_pci_bus_id(device_index), "numa_node": device_numa_node(device_index),
            "local_cpus": len(cpus) if cpus else None, "process_cpus": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None}
