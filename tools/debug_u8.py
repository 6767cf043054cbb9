import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from embclip_b200.encoder import ClipRN50Encoder
from embclip_b200.synthetic import synthetic_rn50_state_dict
enc = ClipRN50Encoder(synthetic_rn50_state_dict(), "cuda:0")
g = torch.Generator().manual_seed(3)
u8 = torch.randint(0, 256, (3, 224, 224, 3), generator=g, dtype=torch.uint8)
mean, std = torch.tensor(enc.CLIP_RGB_MEANS), torch.tensor(enc.CLIP_RGB_STDS)
f32 = (u8.float() / 255.0 - mean) / std
def rel(a, b): return ((a.float() - b.float()).flatten(1).norm(dim=1) / b.float().flatten(1).norm(dim=1)).max().item()
a = enc(u8.cuda(), want=("trunk",)); acts_a = {k: v.clone() for k, v in enc.activations(3).items()}; ta = a["trunk"].clone()
b = enc(f32.cuda(), want=("trunk",)); acts_b = enc.activations(3)
torch.cuda.synchronize()
for k in ["stem.conv1", "stem.conv2", "stem.conv3", "layer1.0.conv3", "layer2.0.conv3", "layer3.0.conv3", "layer4.2.conv3"]:
    print(k, rel(acts_a[k], acts_b[k]), (acts_a[k] != acts_b[k]).float().mean().item())
print("trunk", rel(ta, b["trunk"]))
