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
ta = torch.empty(HEIGHT * WIDTH * 4, dtype=torch.float32).pin_memory()
tb = torch.empty(HEIGHT * WIDTH * 4, dtype=torch.float32).pin_memory()
hosts = [ta.numpy().reshape(HEIGHT, WIDTH, 4), tb.numpy().reshape(HEIGHT, WIDTH, 4)]
rt.set_gi(True, 1)
rt.svo_update(force_full=True); rt.synchronize()

def ev():
    e = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e.record()
    return e

def run(bands, n=8, gi=True):
    rt.set_gi(gi, 1)
    rows = []
    tickets = []
    for i in range(n):
        rt.set_frame_sink(hosts[i % 2], bands)
        a = ev(); rt.clear(); rt.render_visibility(); b = ev(); rt.render_shading(); c = ev()
        rows.append((a, b, c))
        tickets.append(rt.frame_ticket())
    rt.wait_frame(tickets[-1]); rt.synchronize()
    base = rows[0][0]
    print(f"bands={bands} gi={gi}")
    for i, (a, b, c) in enumerate(rows):
        print(f"  frame {i}: start {base.elapsed_time(a):7.3f}  K1 {a.elapsed_time(b):6.3f}  shading {b.elapsed_time(c):6.3f}  end {base.elapsed_time(c):7.3f}")

run(1); run(1); run(2)
rt.destroy()
