"""Host-side placement for the copies either side of the path: run a rank on the CPUs of the NUMA node its GPU hangs
off, so that the pinned host arrays it allocates afterwards are local to that GPU's PCIe root.  With one process per
GPU and no binding, the ranks' host buffers can all land on one socket and every copy of the other socket's GPUs
crosses the inter-socket link: the aggregate host->device rate of 8 ranks then stops near what 2 ranks reach.

    bind_near_gpu(local_rank)      # before the first host allocation

Uses NVML's ideal-CPU mask of the device (matched by UUID, so CUDA_VISIBLE_DEVICES is honoured), else sysfs
(`/sys/bus/pci/devices/<id>/numa_node`).  Never raises: returns None when the topology cannot be read or the
machine has one node.
"""
from __future__ import annotations

import os


def parse_cpulist(text: str):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11]   (the format of /sys/devices/system/node/nodeN/cpulist)"""
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def _cpus_from_nvml(uuid: str):
    import pynvml
    pynvml.nvmlInit()
    try:
        h = pynvml.nvmlDeviceGetHandleByUUID(uuid if uuid.startswith("GPU-") else "GPU-" + uuid)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (int(mask[i // 64]) >> (i % 64)) & 1]
        node = None
        try:
            node = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
        return cpus, node
    finally:
        pynvml.nvmlShutdown()


def _cpus_from_sysfs(domain: int, bus: int, device: int):
    path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/numa_node" % (domain, bus, device)
    node = int(open(path).read())
    if node < 0:
        return [], None
    return parse_cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read()), node


def bind_near_gpu(device_index: int):
    """Restrict the calling thread (and the threads it starts later) to the CPUs next to CUDA device `device_index`.
    Returns {"numa_node", "cpus", "source"} or None if nothing was changed."""
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        cpus, node, source = [], None, None
        try:
            cpus, node = _cpus_from_nvml(str(props.uuid))
            source = "nvml"
        except Exception:
            cpus = []
        if not cpus:
            cpus, node = _cpus_from_sysfs(props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
            source = "sysfs"
        allowed = os.sched_getaffinity(0)
        cpus = sorted(set(cpus) & allowed)
        if not cpus or len(cpus) == len(allowed):
            return None                                  # one node, or a container that already confines us
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus), "source": source}
    except Exception:
        return None
