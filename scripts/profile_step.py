"""One proof of the bench workload bracketed by cudaProfilerStart/Stop, for ncu --profile-from-start off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import hyper_greco_b200
from hyper_greco_b200 import api
name = sys.argv[1] if len(sys.argv) > 1 else bench.DEFAULT_CONFIG
P, inp, bounds, segs, nv = bench.make_case(name, 0)
ctx = api.Context(0)
pp = api.LassoPreprocessing.preprocess(bounds)
node = api.LassoNode(ctx, pp, nv, segs)
d = api.DeviceBuffer.from_numpy(ctx, inp)
for _ in range(2):
    node.prove_claim_reduction(d, api.Keccak256Transcript(), 0, n_inputs=inp.size)
torch.cuda.synchronize()
torch.cuda.profiler.start()
node.prove_claim_reduction(d, api.Keccak256Transcript(), 0, n_inputs=inp.size)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
