"""Per-parameter gradient error of the cfg4-width AFNONet step against the fp64 oracle (the test's setup, all errors printed)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dlwp_benchmark_b200 import fourcastnet as fcn
from oracle import afno_oracle as ao
def rel_l2(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm())
torch.manual_seed(11)
net = fcn.AFNONet(img_height=32, img_width=64, patch_size=(1, 1), constant_channels=4, prescribed_channels=1,
                  prognostic_channels=8, embed_dim=256, depth=2, mlp_ratio=4., num_blocks=8, context_size=1)
with torch.no_grad():
    for n, p in net.named_parameters():
        if ".filter." in n: p.mul_(10.0)
sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
g = torch.Generator().manual_seed(5)
x_t = torch.randn(2, 13, 32, 64, generator=g); gy = torch.randn(2, 8, 32, 64, generator=g)
leaves = {k: v.double().requires_grad_(True) for k, v in sd.items()}
yo = ao.afnonet_step(leaves, x_t.double(), (1, 1), 2, 8); yo.backward(gy.double())
net = net.to("cuda"); y = net.step(x_t.cuda()); y.backward(gy.cuda())
print("y", rel_l2(y, yo))
for k, p in net.named_parameters():
    if leaves[k].grad is not None: print(f"{k:40s} {rel_l2(p.grad, leaves[k].grad):.3e}")
