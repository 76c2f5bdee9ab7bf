"""Do independent traces on different host threads / CUDA streams overlap?  (C5 diagnosis)"""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "parallel-reverb-raytracer_b200", "python")); sys.path.insert(0, ROOT)
import numpy as np, torch
from rayverb_b200 import api, scenes
torch.cuda.set_device(0); api.set_device(0)
tri, vtx, srf = scenes.concert_hall()
dirs = torch.from_numpy(scenes.sphere_directions(1 << 14, seed=1)).cuda()
kw = dict(speakers=[((-1, 0, -1), 0.5), ((1, 0, -1), 0.5)], sr=44100.0, ft="linkwitz_riley", lo_cutoff=45.0)
NJOBS = 16

def worker(k, nworkers, what, rts, streams, out):
    torch.cuda.set_device(0); api.set_device(0); api.set_stream(streams[k].cuda_stream)
    t = []
    for j in range(k, NJOBS, nworkers):
        t0 = time.perf_counter()
        tr = rts[k].raytrace((24, 1.5, 18), (20, 2.0, 4), None, device_ptr=dirs.data_ptr(), nrays=1 << 14,
                             flags=api.TRACE_NO_IMAGES if what == "steady" else 0)
        t1 = time.perf_counter()
        if what == "full":
            tr.render_ir(**kw)
        t.append((1e3 * (t1 - t0), 1e3 * (time.perf_counter() - t1)))
    streams[k].synchronize()
    out[k] = t

for what in ("steady", "trace", "full"):
    for nworkers in (1, 2, 4):
        rts = [api.Raytracer(128, tri, vtx, srf) for _ in range(nworkers)]
        streams = [torch.cuda.Stream() for _ in range(nworkers)]
        for rep in range(2):
            out = [None] * nworkers
            torch.cuda.synchronize(); t0 = time.perf_counter()
            th = [threading.Thread(target=worker, args=(k, nworkers, what, rts, streams, out)) for k in range(nworkers)]
            [x.start() for x in th]; [x.join() for x in th]
            torch.cuda.synchronize(); wall = 1e3 * (time.perf_counter() - t0)
        per = np.array([x for o in out for x in o])
        print(f"{what:6s} workers {nworkers}: {NJOBS} jobs in {wall:7.1f} ms; per job trace {per[:,0].mean():6.2f} ms render {per[:,1].mean():6.2f} ms", flush=True)
        for r in rts: r.close()
