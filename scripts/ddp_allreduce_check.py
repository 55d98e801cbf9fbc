"""Gradient all-reduce of the DDP training loop on real GPUs (NCCL over NVLink): correctness of the in-place arena collective and of the
reference-style entry point, and its time for a detector-sized arena (3.8 M fp32 parameters = 15 MB).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/ddp_allreduce_check.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "se-ssd_b200"))
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from det3d.core.utils import allreduce_grads
from det3d.torchie.trainer import update_ema_variables
from sessd_b200.train import ArenaAdamW, ParamArena, allreduce_grad_arena

torch.manual_seed(0)
model = torch.nn.Sequential(*[torch.nn.Linear(975, 975) for _ in range(4)]).cuda()          # 3.8 M parameters, like the car detector
ema = torch.nn.Sequential(*[torch.nn.Linear(975, 975) for _ in range(4)]).cuda()
arena, arena_ema = ParamArena(model), ParamArena(ema, with_grad=False)
gens = [torch.Generator(device="cuda").manual_seed(100 + r) for r in range(world)]
want = sum(torch.randn(arena.numel, device="cuda", generator=g) for g in gens) / world
arena.grad_flat.copy_(torch.randn(arena.numel, device="cuda", generator=torch.Generator(device="cuda").manual_seed(100 + rank)))
allreduce_grad_arena(arena)
ok1 = bool(torch.allclose(arena.grad_flat, want, rtol=1e-5, atol=1e-6))
arena.grad_flat.copy_(torch.randn(arena.numel, device="cuda", generator=torch.Generator(device="cuda").manual_seed(100 + rank)))
allreduce_grads(model.parameters())                      # reference entry point: finds the arena, one in-place collective
ok2 = bool(torch.allclose(arena.grad_flat, want, rtol=1e-5, atol=1e-6))
# one optimiser-side step: all-reduce -> AdamW -> EMA, timed on the device (max over ranks)
opt = ArenaAdamW(arena)
for _ in range(5):
    allreduce_grad_arena(arena); opt.step(); update_ema_variables(arena, arena_ema, 10)
torch.cuda.synchronize(); dist.barrier()
a, b, c, d = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
iters = 50
a.record()
for _ in range(iters):
    allreduce_grad_arena(arena)
b.record()
for _ in range(iters):
    opt.step()
c.record()
for _ in range(iters):
    update_ema_variables(arena, arena_ema, 10)
d.record()
torch.cuda.synchronize()
t = torch.tensor([a.elapsed_time(b), b.elapsed_time(c), c.elapsed_time(d)], device="cuda") / iters
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    nbytes = arena.numel * 4
    print(json.dumps({"world": world, "params": arena.numel, "arena_allreduce_ok": ok1, "allreduce_grads_ok": ok2,
                      "allreduce_ms": round(float(t[0]), 4), "allreduce_busbw_GBps": round(2 * (world - 1) / world * nbytes / (float(t[0]) / 1e3) / 1e9, 1),
                      "adamw_ms": round(float(t[1]), 4), "adamw_GBps": round(28 * arena.numel / (float(t[1]) / 1e3) / 1e9, 1),
                      "ema_ms": round(float(t[2]), 4), "ema_GBps": round(12 * arena.numel / (float(t[2]) / 1e3) / 1e9, 1)}))
dist.destroy_process_group()
