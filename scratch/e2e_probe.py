import os, sys, time, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tg_b200
from tg_b200.raytracer import from_scene
from bench import build_scene, WIDTH, HEIGHT
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
scene = build_scene(0, 1)
rt = from_scene(scene, device=0)
lib = tg_b200.lib()
stream = torch.cuda.ExternalStream(lib.tgb200_stream(C.byref(rt._rt)), device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ta = torch.empty(HEIGHT * WIDTH * 4, dtype=torch.float32).pin_memory()
tb = torch.empty(HEIGHT * WIDTH * 4, dtype=torch.float32).pin_memory()
print("pinned", ta.is_pinned(), tb.is_pinned())
hosts = [ta.numpy().reshape(HEIGHT, WIDTH, 4), tb.numpy().reshape(HEIGHT, WIDTH, 4)]
rt.set_gi(True, 1)
rt.svo_update(force_full=True); rt.synchronize()

def loop(n, bands, sink=True, wait_each=True, do_flush=True, lag=1):
    tickets = []
    for i in range(n):
        if do_flush:
            with torch.cuda.stream(stream):
                flush.fill_(1)
        if sink:
            rt.set_frame_sink(hosts[i % 2], bands)
        rt.clear(); rt.render()
        if sink:
            tickets.append(rt.frame_ticket())
            if wait_each and i >= lag:
                rt.wait_frame(tickets[i - lag])
    if sink:
        rt.wait_frame(tickets[-1])
    rt.synchronize()

def timeit(name, **kw):
    loop(3, **kw); rt.synchronize()
    t0 = time.perf_counter(); loop(20, **kw); dt = time.perf_counter() - t0
    print(f"{name:50s} {1e3 * dt / 20:.3f} ms/frame", flush=True)

timeit("no sink, flush", bands=1, sink=False)
timeit("no sink, no flush", bands=1, sink=False, do_flush=False)
timeit("sink 8 bands, wait lag 1", bands=8)
timeit("sink 8 bands, wait only at end (unsafe reuse)", bands=8, wait_each=False)
timeit("sink 8 bands, lag 0 (serial frames)", bands=8, lag=0)
timeit("sink 1 band, lag 1", bands=1)
timeit("sink 4 bands, lag 1", bands=4)
timeit("sink 2 bands, lag 1", bands=2)
timeit("sink 3 bands, lag 1", bands=3)
timeit("sink 8 bands, no flush", bands=8, do_flush=False)
# submission cost on the host: time to enqueue one frame
rt.synchronize()
rt.set_frame_sink(hosts[0], 8)
t0 = time.perf_counter(); rt.clear(); rt.render(); t1 = time.perf_counter(); rt.synchronize()
print(f"host time to enqueue one frame with 8 bands: {1e3 * (t1 - t0):.3f} ms")
rt.set_frame_sink(None)
t0 = time.perf_counter(); rt.clear(); rt.render(); t1 = time.perf_counter(); rt.synchronize()
print(f"host time to enqueue one frame without sink: {1e3 * (t1 - t0):.3f} ms")
rt.destroy()
