"""GPU diagnostic: product model (CUDA kernels) vs oracle model (CPU fp32) stage by stage."""
import os, sys, time, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from unittest import mock
import eosvos_b200
from eosvos_b200.networks.mask_rcnn import MaskRCNN
from eosvos_b200.meta_optim.meta_optim import MetaOptimizer
from oracle import model_oracle as MO

dev = torch.device("cuda:0")
_real_randperm = torch.randperm

def det_randperm(gen):
    def f(n, *a, device=None, **k):
        return _real_randperm(n, generator=gen).to(device if device is not None else "cpu")
    return f

def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()

def nchw(x): return x.float().permute(0, 3, 1, 2)

def main(min_size=160, max_size=266, h=96, w=170, kind="LOVASZ"):
    g = torch.Generator().manual_seed(11)
    img = torch.rand(1, 3, h, w, generator=g)
    tgt = torch.zeros(1, 1, h, w); tgt[0, 0, int(h*0.3):int(h*0.73), int(w*0.3):int(w*0.7)] = 1
    oracle = MO.build_oracle_model(seed=1, maskrcnn_loss=kind, min_size=min_size, max_size=max_size)
    torch.manual_seed(1)
    model = MaskRCNN('resnet50', num_classes=2, batch_norm={'accum_stats': False, 'learn_weight': False, 'learn_bias': False},
                     train_encoder=True, roi_pool_output_sizes={'box': 7, 'mask': 28}, eval_augment_rpn_proposals_mode='EXTEND',
                     replace_batch_with_group_norms=True, box_nms_thresh=0.5, maskrcnn_loss=kind)
    if min_size is not None:
        model.transform.min_size = (min_size,); model.transform.max_size = max_size
    torch.manual_seed(3); oopt = MO.OracleMetaOptimizer(oracle, 1e-3); oopt.reset()
    torch.manual_seed(3); opt = MetaOptimizer(model, 1e-3, True, False, 'NEURON', False, None)
    model.to(dev); opt.to(dev); opt.reset(); opt.eval()
    oracle.train_without_dropout(); model.train_without_dropout()

    # ---------------- train forward, oracle first
    with mock.patch("torch.randperm", det_randperm(torch.Generator().manual_seed(5))):
        torch.manual_seed(21); t0 = time.time(); oloss, olosses = oracle(img, tgt); t_or = time.time() - t0
    model.capture = {}
    model.fixed_proposals = oracle.last_proposals
    with mock.patch("torch.randperm", det_randperm(torch.Generator().manual_seed(5))):
        torch.manual_seed(21); loss, losses = model(img.to(dev), tgt.to(dev))
    torch.cuda.synchronize()
    cap = model.capture
    print(f"oracle train fwd {t_or:.2f}s; loss oracle {oloss.item():.5f} product {loss.item():.5f}")
    for k in olosses: print(f"   {k}: oracle {olosses[k].item():.6f} product {losses[k].item():.6f}")
    ofe = list(oracle.last_features.values())
    print("   transform rel:", rel(nchw(cap["x8"])[:, :3], oracle.last_images.tensors))
    for i, (a, b) in enumerate(zip(cap["feats"], ofe)): print(f"   P{i+2} rel: {rel(nchw(a), b):.4f}  shape {tuple(b.shape)}")
    print("   rpn objectness rel:", rel(cap["objectness"], oracle.last_rpn_raw[0]), " deltas rel:", rel(cap["deltas"], oracle.last_rpn_raw[1]))
    same_samples = all(torch.equal(a.cpu(), b) for a, b in zip(cap["sampled_proposals"], oracle.last_sampled_proposals))
    print("   same sampled proposals:", same_samples)
    if same_samples:
        print("   class_logits rel:", rel(cap["class_logits"], oracle.last_box_raw[0]), " box_reg rel:", rel(cap["box_regression"], oracle.last_box_raw[1]))
        if "mask_logits" in cap and cap["mask_logits"].shape == oracle.last_mask_logits.shape:
            print("   mask_logits rel:", rel(cap["mask_logits"], oracle.last_mask_logits), tuple(cap["mask_logits"].shape))
    # ---------------- grads + step
    t0 = time.time(); ogr = oopt.step(oloss); t_ob = time.time() - t0
    groups = list(opt.meta_model.param_groups())
    params = [p for *_, p in groups]
    grads = torch.autograd.grad(loss, params, retain_graph=True)
    names = [f"{a}.{c}" for a, _, c, _ in groups]
    errs = sorted(((rel(gp, go), n, go.norm().item()) for n, gp, go in zip(names, grads, ogr)), reverse=True)
    print(f"oracle bwd {t_ob:.2f}s; grad rel err: median {errs[len(errs)//2][0]:.4f}  worst:")
    for e in errs[:12]: print(f"      {e[0]:.4f}  {e[1]}  |g|={e[2]:.3e}")
    import collections
    proj = collections.defaultdict(lambda: [0.0, 0.0])
    for n, gp, go in zip(names, grads, ogr):
        kind = ("bias" if n.endswith(".bias") else "weight") + ("/gn" if ".bn" in n or "downsample.1" in n else "") + "/" + n.split(".")[0] + ("." + n.split(".")[1] if n.startswith("roi_heads") or n.startswith("backbone") else "")
        proj[kind][0] += (gp.double().cpu() * go.double()).sum().item(); proj[kind][1] += (go.double() ** 2).sum().item()
    print("   projection <g,g_ref>/<g_ref,g_ref> per group:", {k: round(a / max(b, 1e-30), 3) for k, (a, b) in proj.items()})
    tot = torch.sqrt(sum(((gp.double().cpu() - go.double()) ** 2).sum() for gp, go in zip(grads, ogr))) / torch.sqrt(sum((go.double() ** 2).sum() for go in ogr))
    print("   global grad rel err:", tot.item())
    opt.set_train_loss(loss); opt.step(loss); opt.meta_model.detach_param_groups()
    perr = max(rel(p, q) for (*_, p), (*_, q) in zip(opt.meta_model.param_groups(), oopt.groups()))
    print("   max param rel err after step:", perr)
    # ---------------- eval forward
    model.fixed_proposals = None; oracle.fixed_proposals = None; model.capture = None
    oracle.eval(); model.eval()
    torch.manual_seed(22)
    with torch.no_grad(): oprobs, oboxes = oracle(img, tgt)
    torch.manual_seed(22)
    with torch.no_grad(): probs, boxes = model(img.to(dev), tgt.to(dev))
    print("eval boxes oracle", oboxes.flatten().tolist(), "product", boxes.flatten().tolist())
    pm, om = probs.cpu() >= 0.5, oprobs >= 0.5
    iou = (pm & om).sum().item() / max((pm | om).sum().item(), 1)
    print(f"eval mask IoU {iou:.4f}; max prob diff {(probs.cpu()-oprobs).abs().max().item():.4f}; fg px oracle {om.sum().item()} product {pm.sum().item()}")

if __name__ == "__main__":
    full = len(sys.argv) > 1 and sys.argv[1] == "full"
    if full: main(None, None, 480, 854)
    else: main()
