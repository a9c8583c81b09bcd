#!/usr/bin/env python
"""Developer probe: NUMA layout of the box, which CPUs/memory nodes this process may use, GPU -> node mapping, and whether the
memory-policy system calls (mbind, set_mempolicy, move_pages) are permitted inside this container."""
import ctypes, glob, os, subprocess
print("nodes:", sorted(os.path.basename(p) for p in glob.glob("/sys/devices/system/node/node[0-9]*")))
for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    try:
        print(os.path.basename(p), "cpulist", open(p + "/cpulist").read().strip(), "| MemTotal", [l for l in open(p + "/meminfo") if "MemTotal" in l][0].split()[-2], "kB")
    except OSError as e:
        print(p, e)
for l in open("/proc/self/status"):
    if l.startswith(("Cpus_allowed_list", "Mems_allowed_list")):
        print(l.strip())
print("sched_getaffinity:", sorted(os.sched_getaffinity(0)))
try:
    print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
    out = subprocess.run(["nvidia-smi", "--query-gpu=index,pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout
    for l in out.strip().splitlines():
        idx, bus = [x.strip() for x in l.split(",")]
        b = bus.lower()
        if b.startswith("00000000:"):
            b = b[4:]
        try:
            print("gpu", idx, bus, "numa_node", open("/sys/bus/pci/devices/%s/numa_node" % b).read().strip())
        except OSError as e:
            print("gpu", idx, bus, e)
except OSError as e:
    print(e)
libc = ctypes.CDLL(None, use_errno=True)
SYS_mbind, SYS_set_mempolicy, SYS_get_mempolicy = 237, 238, 239
import mmap
m = mmap.mmap(-1, 1 << 22)
addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
mask = ctypes.c_ulong(1)
for name, args in (("mbind(MPOL_BIND node0)", (SYS_mbind, ctypes.c_void_p(addr), ctypes.c_ulong(1 << 22), 2, ctypes.byref(mask), ctypes.c_ulong(64), 0)),
                   ("mbind(MPOL_PREFERRED node0)", (SYS_mbind, ctypes.c_void_p(addr), ctypes.c_ulong(1 << 22), 1, ctypes.byref(mask), ctypes.c_ulong(64), 0)),
                   ("set_mempolicy(MPOL_DEFAULT)", (SYS_set_mempolicy, 0, None, ctypes.c_ulong(0)))):
    ctypes.set_errno(0)
    rc = libc.syscall(*args)
    print(name, "rc", rc, "errno", ctypes.get_errno(), os.strerror(ctypes.get_errno()) if rc else "")
