"""Host placement for the streaming path: run next to the GPU's PCIe root.

`ChannelBank.stream` is bound by the host-to-device copy of the raw capture (52 GB/s per GPU on this box).  With
several ranks on one node the pinned buffers and the reader threads must sit on the NUMA node the GPU hangs off:
round 1 ran all eight ranks on node 0 and the box delivered 172 GB/s in aggregate instead of 8 x 52.
`bind_to_device_numa` pins the calling process to that node's CPUs; page-locked buffers allocated afterwards
(`iq2a_host_alloc`, first touch under the default local policy) then land in its memory.
"""
from __future__ import annotations

import os
from pathlib import Path


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def device_numa_node(index: int) -> int | None:
    """NUMA node of CUDA device `index` (sysfs, through the PCI bus id NVML reports); None when unknown."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
    except Exception:
        return None
    # NVML: 00000000:1B:00.0 -> sysfs: 0000:1b:00.0
    dom, rest = bus.split(":", 1)
    path = Path("/sys/bus/pci/devices") / f"{dom[-4:]}:{rest}".lower() / "numa_node"
    try:
        node = int(path.read_text().strip())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def device_cpu_affinity(index: int) -> set[int] | None:
    """The CPUs NVML calls ideal for device `index` (the fallback when sysfs has no NUMA node for the PCI device,
    as in some virtualised boxes); None when NVML cannot say."""
    try:
        import pynvml
        pynvml.nvmlInit()
        ncpu = max(os.cpu_count() or 1, 64)
        words = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(index), (ncpu + 63) // 64)
    except Exception:
        return None
    cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
    return cpus or None


def host_topology() -> dict:
    """NUMA nodes the kernel exposes and the CPUs this process may use (for the logs next to a throughput figure)."""
    try:
        nodes = sorted(int(p.name[4:]) for p in Path("/sys/devices/system/node").glob("node[0-9]*"))
    except OSError:
        nodes = []
    try:
        allowed = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        allowed = os.cpu_count() or 0
    return {"numa_nodes": len(nodes), "cpus_allowed": allowed}


def bind_to_device_numa(index: int) -> dict:
    """Restrict the calling process to the CPUs of the GPU's NUMA node (no-op when the topology is unknown, the node
    has no CPUs in this cgroup, or the platform has no sched_setaffinity).  Returns what was done, for the logs."""
    info = {"device": index, "numa_node": device_numa_node(index), "bound": False}
    node = info["numa_node"]
    info.update(host_topology())
    if not hasattr(os, "sched_setaffinity"):
        return info
    if node is None:
        # no node in sysfs: take NVML's ideal CPU set when it is a proper subset of what we may use
        ideal = device_cpu_affinity(index)
        try:
            allowed = os.sched_getaffinity(0)
            use = sorted((ideal or set()) & allowed)
            if use and len(use) < len(allowed):
                os.sched_setaffinity(0, use)
                info.update(bound=True, cpus=len(use), method="nvml cpu affinity")
        except OSError:
            pass
        return info
    try:
        node_cpus = _parse_cpulist((Path("/sys/devices/system/node") / f"node{node}" / "cpulist").read_text())
        allowed = os.sched_getaffinity(0)
        use = sorted(node_cpus & allowed)
        if use:
            os.sched_setaffinity(0, use)
            info.update(bound=True, cpus=len(use))
    except OSError:
        pass
    return info
