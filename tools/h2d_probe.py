#!/usr/bin/env python3
"""Host-to-device copy rate from page-locked memory of three kinds: torch's pinned allocator, cudaHostAlloc default, and
cudaHostAlloc write-combined (CUDA events around 20 copies of 128 MB on one stream).
    python tools/h2d_probe.py"""
import ctypes
import json

import torch


def main():
    torch.cuda.set_device(0)
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
    rt.cudaFreeHost.argtypes = [ctypes.c_void_p]
    rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
    n = 128 << 20
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream()
    out = {}

    def rate(src_ptr):
        for _ in range(3):
            rt.cudaMemcpyAsync(dev.data_ptr(), src_ptr, n, 1, st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            rt.cudaMemcpyAsync(dev.data_ptr(), src_ptr, n, 1, st.cuda_stream)
        e1.record()
        torch.cuda.synchronize()
        return 20 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9

    pinned = torch.empty(n, dtype=torch.uint8).pin_memory()
    pinned.fill_(7)
    direct = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    direct.fill_(7)
    ptrs = {"torch_pin_memory()": pinned.data_ptr(), "torch_empty(pin_memory=True)": direct.data_ptr()}
    raw = []
    for name, flag in (("cudaHostAlloc_default", 0), ("cudaHostAlloc_write_combined", 4)):
        p = ctypes.c_void_p()
        if rt.cudaHostAlloc(ctypes.byref(p), n, flag) == 0:
            ctypes.memset(p, 7, n)
            ptrs[name] = p.value
            raw.append(p)
    for rnd in range(3):  # interleaved rounds: the first copies of a process also wake the link up
        for name, ptr in ptrs.items():
            out.setdefault(name, []).append(round(rate(ptr), 2))
    for p in raw:
        rt.cudaFreeHost(p)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
