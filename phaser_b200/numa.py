"""Host-side placement of a rank next to its GPU.

The end-to-end rate of the path is the host -> device copy of the packed records.  On a
two-socket box a page-locked buffer that lives on the other socket crosses the inter-socket link on its way to the
GPU, and eight ranks started by torchrun all allocate wherever the kernel first schedules them.  bind_near_gpu()
pins the calling process to the cores of the NUMA node the GPU's PCIe root hangs off BEFORE the transport buffers are
allocated and filled, so that first-touch puts them on that node.  Pure sysfs, no libnuma; a box that does not expose
the topology (node -1, a container without /sys/devices/system/node) is left alone and the fact is reported.
"""
import os


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(text):
    cpus = set()
    for part in (text or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(device_index):
    """NUMA node of CUDA device `device_index` (None when the box does not say)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = "%04x:%02x:%02x.0" % (getattr(p, "pci_domain_id", 0), p.pci_bus_id, p.pci_device_id)
    except Exception:
        return None
    node = _read("/sys/bus/pci/devices/%s/numa_node" % bdf)
    try:
        node = int(node)
    except (TypeError, ValueError):
        return None
    return node if node >= 0 else None


def bind_near_gpu(device_index):
    """Returns a dict describing what was done: {"numa_node": n or None, "cpus": count, "bound": bool, "nodes": total}."""
    nodes = [d for d in (os.listdir("/sys/devices/system/node") if os.path.isdir("/sys/devices/system/node") else [])
             if d.startswith("node") and d[4:].isdigit()]
    info = {"numa_node": gpu_numa_node(device_index), "nodes": len(nodes), "bound": False, "cpus": len(os.sched_getaffinity(0))}
    if info["numa_node"] is None or len(nodes) < 2:
        return info
    cpus = _parse_cpulist(_read("/sys/devices/system/node/node%d/cpulist" % info["numa_node"]))
    cpus &= os.sched_getaffinity(0)
    if not cpus:
        return info
    try:
        os.sched_setaffinity(0, cpus)
        info["bound"] = True; info["cpus"] = len(cpus)
    except OSError:
        pass
    return info
