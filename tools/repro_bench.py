import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
import torch, numpy as np
sys.argv = sys.argv  # noqa
import bench
S = int(sys.argv[1]) if len(sys.argv) > 1 else 64
use_default_stream = (sys.argv[2] == "default") if len(sys.argv) > 2 else True
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 8
capi = bench.load_mod("sdrb_capi", "real-time-sdr_b200/capi.py")
gen = bench.load_mod("sdrgen", "real-time-sdr_b200/sdrgen.py")
dev = torch.device("cuda", 0)
ch = capi.Chain(0, "r", n_streams=S)
bb = ch.info.block_bytes
pitch = (bb + 255) // 256 * 256
if use_default_stream:
    ch.set_stream(torch.cuda.current_stream().cuda_stream)
inputs = bench.build_inputs(torch, gen, S, bb, pitch, dev)
torch.cuda.synchronize()
for i in range(nsteps):
    ch.process_device(inputs[i % len(inputs)].data_ptr(), pitch)
    if len(sys.argv) <= 4 or sys.argv[4] != "nosync":
        ch.sync()
        torch.cuda.synchronize()
        print("step", i, "ok", flush=True)
ch.sync()
torch.cuda.synchronize()
print("done")
