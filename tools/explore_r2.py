"""Round-2 exploration on the GPU box: (1) the UNMODIFIED reference (oracle/_ref) on CUDA through cuDNN/ATen -- time per
fine-tune iteration (batch 3) and per inference frame at 854x480; (2) n_pos / n_det of the product in the bench block."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def ref_gpu(tf32):
    from oracle import ref_harness as RH, ref_shims
    RH.install_full()
    import meta_optim.meta_optim as mo
    import bench
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.allow_tf32 = tf32
    dev = torch.device("cuda:0")
    model = ref_shims.build_reference_model(seed=1)
    opt = mo.MetaOptimizer(model, init_lr=1e-3, learn_model_init=True, second_order_gradients=False,
                           lr_hierarchy_level='NEURON', use_log_init_lr=False, max_lr=None)
    model.to(dev)
    opt.to(dev)
    opt.reset()
    opt.eval()
    model.roi_heads.detections_per_img = 1
    fr, gt0, batches = bench.build_workload(1)
    db = [(a.to(dev), b.to(dev)) for a, b in batches]
    model.train_without_dropout()

    def it(i):
        loss, _ = model(*db[i % len(db)])
        model.zero_grad()
        opt.set_train_loss(loss)
        opt.step(loss)
        opt.meta_model.detach_param_groups()
        return loss
    for i in range(3):
        it(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for i in range(n):
        it(i)
    torch.cuda.synchronize()
    t_it = (time.perf_counter() - t0) / n
    model.eval()
    tgt = gt0[None, None].to(dev)
    frames = [fr[1 + i:2 + i].to(dev) for i in range(3)]
    with torch.no_grad(), RH.reference_on_device(dev):
        for f in frames:
            model(f, tgt)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            for f in frames:
                p, b = model(f, tgt)
        torch.cuda.synchronize()
    t_fr = (time.perf_counter() - t0) / 9
    print(f"reference on CUDA (tf32={tf32}): {1 / t_it:.2f} iter/s ({t_it * 1e3:.1f} ms), {1 / t_fr:.2f} frames/s "
          f"({t_fr * 1e3:.1f} ms); boxes {b.flatten().tolist()}", flush=True)


def product_counts():
    import bench
    from eosvos_b200.util import evaluate as E
    dev = torch.device("cuda:0")
    model, opt = bench.build_model(dev)
    fr, gt0, batches = bench.build_workload(1)
    db = [(a.to(dev), b.to(dev)) for a, b in batches]
    tgt = gt0[None, None].to(dev)
    for thr in (0.5, 0.05):
        model.roi_heads.score_thresh = thr
        opt.reset()
        for iters in (10, 20, 40):
            E.finetune(model, opt, lambda e: db[e % len(db)], 10, 1, 1)
            model.eval()
            model.capture = {}
            with torch.no_grad():
                p, b = model(fr[1:2].to(dev), tgt)
            det = model.capture["detections"][0]
            model.capture = None
            print(f"score_thresh {thr} after {iters} iters: n_det {det['boxes'].shape[0]} scores {det['scores'].tolist()} "
                  f"mask px {(p >= 0.5).sum().item()}", flush=True)


if __name__ == "__main__":
    product_counts()
    ref_gpu(False)
    ref_gpu(True)
