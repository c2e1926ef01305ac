"""Compares engine intermediates around the stems with a torch fp64 autograd reference (run on the GPU box)."""
import torch
import torch.nn.functional as F
from oracle import torch_oracle as O
from radar_depth_b200.model.models import ResNet_latefusion
from radar_depth_b200.evaluation.criteria_new import MaskedL1Loss

h, w, cin = 64, 96, 4
sd = O.synth_state_dict(O.latefusion_entries(cin))
inputs, target = O.synth_batch(2, h, w)
params, work = O._leaf_params(sd, torch.float64)
x = inputs.double()
nb = {}
xi = O._conv(x[:, :3], work, "conv1", 2, 3); zi = xi; zi.retain_grad()
xi = F.relu(O._bn(xi, work, "bn1", True, nb))
pi = F.max_pool2d(xi, 3, 2, 1); pi.retain_grad()
ei = O._encoder(pi, work, "", True, nb)
xd = O._conv(x[:, 3:], work, "conv1_depth", 2, 3); zd = xd; zd.retain_grad()
xd = F.leaky_relu(O._bn(xd, work, "bn1_depth", True, nb), 0.2)
pd = F.max_pool2d(xd, 3, 2, 1); pd.retain_grad()
ed = O._encoder(pd, work, "_depth", True, nb)
f = torch.cat((ei, ed), 1); f.retain_grad()
f2 = O._bn(O._conv(f, work, "conv_fusion"), work, "bn_fusion", True, nb)
f2 = O._bn(O._conv(f2, work, "conv2"), work, "bn2", True, nb)
for li in range(1, 5):
    f2 = O._upproj(f2, work, f"decoder.layer{li}", True, nb)
c3 = O._conv(f2, work, "conv3", 1, 1)
pred = F.interpolate(c3, size=(h, w), mode="bilinear", align_corners=True)
loss = O.masked_l1(pred, target.double())
loss.backward()

m = ResNet_latefusion(18, "upproj", (h, w), cin, pretrained=False)
m.load_state_dict(sd, strict=True)
m = m.cuda().train()
m.precision = "fp32"
p = m(inputs.cuda())
l = MaskedL1Loss()(p, target.cuda())
l.backward()
torch.cuda.synchronize()
eng = m._engine
d = eng.dbg


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def nhwc(t):
    return t.permute(0, 2, 3, 1)


print("fwd z_stem rgb", rel(d["z_stem"][..., :64], nhwc(zi.detach())), "depth", rel(d["z_stem"][..., 64:], nhwc(zd.detach())))
print("fwd pooled rgb", rel(d["p_rgb"], nhwc(pi.detach())), "depth", rel(d["p_d"], nhwc(pd.detach())))
print("fwd concat", rel(d["concat"], nhwc(f.detach())))
print("bwd d_concat", rel(d["d_concat"], nhwc(f.grad)), "rgb part", rel(d["d_concat"][..., :512], nhwc(f.grad)[..., :512]), "depth part", rel(d["d_concat"][..., 512:], nhwc(f.grad)[..., 512:]))
print("bwd dpool rgb", rel(eng.blocks_all[0][0]["dx_t"], nhwc(pi.grad)), "depth", rel(eng.blocks_all[1][0]["dx_t"], nhwc(pd.grad)))
print("bwd gz_stem (dz of stems) rgb", rel(d["gz_stem"][..., :64], nhwc(zi.grad)), "depth", rel(d["gz_stem"][..., 64:], nhwc(zd.grad)))
for Bk in eng.blocks_all[0][:1] + eng.blocks_all[1][:1]:
    print("block", Bk["pfx"])
print("keys", list(d.keys()))
print("dpool rgb ref norm", float(pi.grad.norm()), "depth", float(pd.grad.norm()))
eng.dbg_dpool = None

# ---- isolate the stem parameter gradients
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
named = dict(m.named_parameters())
for k in ("conv1.weight", "bn1.weight", "bn1.bias", "conv1_depth.weight", "bn1_depth.weight", "bn1_depth.bias"):
    print(k, "engine vs oracle-param-grad", rel(named[k].grad, params[k].grad))
gz = d["gz_stem"].float()
xr = inputs.cuda()
w1 = named["conv1.weight"].detach().clone().requires_grad_(True)
y1 = F.conv2d(xr[:, :3], w1, None, 2, 3)
gw1, = torch.autograd.grad(y1, w1, grad_outputs=gz[..., :64].permute(0, 3, 1, 2).contiguous())
print("conv1.weight: engine vs torch-wgrad(engine dz)", rel(named["conv1.weight"].grad, gw1), "torch-wgrad(engine dz) vs oracle", rel(gw1, params["conv1.weight"].grad))
wd = named["conv1_depth.weight"].detach().clone().requires_grad_(True)
yd = F.conv2d(xr[:, 3:], wd, None, 2, 3)
gwd, = torch.autograd.grad(yd, wd, grad_outputs=gz[..., 64:].permute(0, 3, 1, 2).contiguous())
print("conv1_depth.weight: engine vs torch-wgrad(engine dz)", rel(named["conv1_depth.weight"].grad, gwd), "vs oracle", rel(gwd, params["conv1_depth.weight"].grad))
print("engine conv1.weight.grad[0,0]:\n", named["conv1.weight"].grad[0, 0].cpu().numpy())
print("expected:\n", gw1[0, 0].cpu().numpy())
gs = d["g_stem"]
print("bstats sum_g[:4]", gs.bstats[0][:4].cpu().numpy(), "bn1.bias.grad[:4]", named["bn1.bias"].grad[:4].cpu().numpy(), "oracle", params["bn1.bias"].grad[:4].numpy())
