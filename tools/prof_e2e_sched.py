import sys, torch
sys.path.insert(0, '.')
from fbanet_b200 import BaseModel
dev = torch.device('cuda:0')
m = BaseModel(num_frames=14, img_size=160, in_channels=3, embed_dim=64, window_length=10, token_projection="linear", token_mlp="leff", dtype="bf16", seed=0).to(dev).eval()
B = 64
hin = torch.rand(B, 14, 3, 160, 160).pin_memory()
hout = torch.empty(B, 3, 640, 640).pin_memory()
for sched in (32, (16, 32, 16), (8, 24, 24, 8), (16, 16, 16, 16), (24, 24, 16), (16, 24, 24), 64, 32):
    for _ in range(2):
        m.infer_host(hin, hout, chunk=sched)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m.infer_host(hin, hout, chunk=sched)
    e1.record()
    torch.cuda.synchronize()
    print(sched, round(e0.elapsed_time(e1) / 5, 2), "ms", flush=True)
