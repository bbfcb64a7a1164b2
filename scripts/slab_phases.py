"""Per-phase device and host times of the multi-GPU step (run under torchrun)."""
import importlib, os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = importlib.import_module("sph-erosion_b200"); slabs = importlib.import_module("sph-erosion_b200.slabs")
n_axis = int(os.environ.get("NAXIS", "100")); lag = int(os.environ.get("LAG", "0"))
pos, ids, box, bounds = slabs.channel_block(n_axis, world, rank, False)
gy = bench.scene_gravity(n_axis)
layer = int(n_axis * n_axis * (0.0457 * 1.001 / 0.025 + 1)); cap = max(4 * layer, 1 << 14)
sim, b, cols = slabs.make_gpu_slab(pkg, local, rank, world, box, dict(len=box[1], dt=0.01, g=(0.0, gy, 0.0)), bounds, cap)
sim.slab_upload(pos, np.zeros_like(pos), ids)
c = slabs.TorchComm(rank, world)
drv = slabs.SlabDriver(b, c, lag=lag)
for _ in range(10): drv.step()
drv.drain()
K = 50
ev = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(K)]
host = np.zeros((K, 4))
torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(K):
    e = ev[k]
    e[0].record(); h0 = time.perf_counter()
    b.pack(); e[1].record(); h1 = time.perf_counter()
    c.swap_records(b.send_l, drv.out_l, b.send_r, drv.out_r, b.recv_l, drv.in_l, b.recv_r, drv.in_r); e[2].record(); h2 = time.perf_counter()
    if lag: b.unpack_async(b.recv_l, drv.in_l, b.recv_r, drv.in_r)
    else: b.unpack(b.recv_l, drv.in_l, b.recv_r, drv.in_r)
    e[3].record(); h3 = time.perf_counter()
    b.step(); e[4].record(); h4 = time.perf_counter()
    host[k] = [h1 - h0, h2 - h1, h3 - h2, h4 - h3]
torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / K
dev = np.array([[e[i].elapsed_time(e[i + 1]) for i in range(4)] for e in ev])
gap = np.array([ev[k][4].elapsed_time(ev[k + 1][0]) for k in range(K - 1)])
print("rank %d lag %d wall/step %.3f ms | device ms: pack %.3f nccl %.3f unpack %.3f step %.3f gap %.3f | host ms: pack %.3f nccl %.3f unpack %.3f step %.3f"
      % (rank, lag, wall * 1e3, *dev.mean(0), gap.mean(), *(host.mean(0) * 1e3)), flush=True)
dist.barrier(); dist.destroy_process_group()
