"""Multi-GPU probe (run under torchrun): times the all-gathers alone and prints per-rank kernel spans + clocks."""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import chain_harness as ch
from realism_effects_b200 import abi, engine, parallel, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
W, Hr = 3840, 2160; H = Hr * world
o = ch.Opts(denoise_iterations=2)
class I: width, height = W, H
ctx = engine.Context(local)
env = synth.synthetic_env(1024, 512); marg, cond, total = synth.build_env_cdf(env.astype(np.float32)); ctx.set_env(env, marg, cond, total)
chain = parallel.ShardedSsgiChain(ctx, ch.chain_options(I, o), blocks_per_rank=4, overlap=True)
# --- pure exchange timing
s = chain.stream
for rep in range(3):
    with torch.cuda.stream(s):
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(s); chain._gather(chain.plan.gathered_planes[:1]); chain._wait([0]); e1.record(s)
        chain._gather(chain.plan.gathered_planes[1:]); chain._wait([4, 5]); e2.record(s)
    s.synchronize(); dist.barrier()
t_comp, t_dn = e0.elapsed_time(e1), e1.elapsed_time(e2)
recv = chain.exchange_bytes_per_frame
print(f"rank {rank}: all-gather composed {t_comp:.3f} ms, dnB x2 {t_dn:.3f} ms; recv/frame {recv/1e6:.0f} MB => {recv/1e6/(t_comp+t_dn):.0f} GB/s in", flush=True)
q = subprocess.run(["nvidia-smi", "--query-gpu=index,clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader", "-i", str(local)], capture_output=True, text=True).stdout.strip()
print(f"rank {rank} idle clocks: {q}", flush=True)
chain.close(); ctx.close(); dist.destroy_process_group()
