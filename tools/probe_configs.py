"""GPU probe: one train step of every BASELINE.json config on one B200 (clips/s, CUDA events, synthetic data).
cfg 1: notebook small-CNN LRCN 20 x 64x64, 50 classes (fp32 parity path and bf16 gate GEMM), B = 8 and 64
cfg 2: medsos ResNet-50 LRCN at T = 30 (the bench runs T = 16)
cfg 3: ucf50-lrcn topology, frozen ResNet-50 at 224x224 x 16 frames, 4-layer biLSTM H = 56, B = 32"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import video_classif_b200 as vc

dev = torch.device("cuda", 0)


def rate(model, x, y, steps=10, graph=False):
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4, fused=True)
    if graph:
        model.enable_encoder_graph()

    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(model(x), y)
        loss.backward()
        opt.step()
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return x.shape[0] / ms * 1e3, ms


print(torch.cuda.get_device_name(0))
torch.manual_seed(0)
ONLY_GRAPH = "--graph" in sys.argv
for B in (() if ONLY_GRAPH else (8, 64)):
    for prec in ("fp32", "bf16"):
        m = vc.SmallCNNLRCN(50, 20, 32, (3, 64, 64), dropout=0.5, precision=prec).to(dev).train()
        x = torch.rand(B, 20, 3, 64, 64, device=dev) * 255
        y = torch.randint(0, 50, (B,), device=dev)
        r, ms = rate(m, x, y)
        print(f"cfg1 small-CNN LRCN 20x64x64 B={B} {prec}: {r:9.0f} clips/s ({ms:.2f} ms/step)")
if not ONLY_GRAPH:
    m = vc.LRCN(4, 30, 32, 8, cnn_backbone="resnet50", rnn_layers=3, dropout=0.25).to(dev).train()
    x = torch.rand(64, 30, 3, 112, 112, device=dev); y = torch.randint(0, 4, (64,), device=dev)
    r, ms = rate(m, x, y, steps=6, graph=True)
    print(f"cfg2 medsos ResNet-50 LRCN 30x112x112 B=64 bf16: {r:9.0f} clips/s ({ms:.2f} ms/step)")
    del m, x
    torch.cuda.empty_cache()
    m = vc.UCF50LRCN(50, 16, 56, 512, cnn_backbone="resnet50", rnn_layers=4).to(dev).train()
    x = torch.rand(32, 16, 3, 224, 224, device=dev); y = torch.randint(0, 50, (32,), device=dev)
    r, ms = rate(m, x, y, steps=6, graph=True)
    print(f"cfg3 ucf50 ResNet-50 @224 16 frames biLSTM H=56 x4 B=32 bf16: {r:9.0f} clips/s ({ms:.2f} ms/step, "
          f"{130.8 * 32 / ms / 1e0:.0f} GFLOP/ms encoder)")
    del m, x
    torch.cuda.empty_cache()
    # cfg 2, second topology: the crime / rgb scripts' default -- densenet121 fully trainable (FINETUNE = True), one adapt,
    # 4-layer biLSTM H = 56, 3 binary heads (sum of BCE-with-logits), 16 x 112x112, B = 64
    if "--finetune" in sys.argv or True:
        for arch, B in (("densenet121", 64), ("resnet50", 64)):
            m = vc.CrimeLRCN(3, 16, 56, 512, cnn_backbone=arch, finetune=True, rnn_layers=4, classif_mode="multiple_binary").to(dev).train()
            x = torch.rand(B, 16, 3, 112, 112, device=dev)
            yb = (torch.rand(B, 3, device=dev) > 0.5).float()
            opt = torch.optim.Adam(m.parameters(), lr=1e-4, fused=True)

            def step():
                opt.zero_grad(set_to_none=True)
                loss = torch.nn.functional.binary_cross_entropy_with_logits(m(x), yb)
                loss.backward()
                opt.step()
            for _ in range(3):
                step()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"cfg2b crime LRCN {arch} FULL fine-tune 16x112x112 B={B} biLSTM H=56 x4, 3 binary heads: {B / ms * 1e3:9.0f} clips/s "
                  f"({ms:.2f} ms/step, {torch.cuda.max_memory_allocated() / 2**30:.1f} GB)")
            del m, x, opt
            torch.cuda.empty_cache()

# ---- eager loop vs the whole step as one CUDA-graph replay (GraphedTrainStep) at the reference's own small batches
if "--graph" in sys.argv:
    def both(name, make, x, y, crit, steps=10):
        res = []
        for graphed in (False, True):
            torch.manual_seed(0)
            m = make().to(dev).train()
            params = [p for p in m.parameters() if p.requires_grad]
            opt = torch.optim.Adam(params, lr=1e-4, fused=True, capturable=graphed)
            if graphed:
                gs = vc.GraphedTrainStep(m, opt, crit, x, y)
                step = lambda: gs(x, y)
            else:
                def step():
                    opt.zero_grad(set_to_none=True)
                    crit(m(x), y).backward()
                    opt.step()
            for _ in range(3):
                step()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(steps):
                step()
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / steps)
            del m, opt
            torch.cuda.empty_cache()
        B = x.shape[0]
        print(f"{name}: eager {res[0]:7.3f} ms/step ({B / res[0] * 1e3:8.0f} clips/s)   graph replay {res[1]:7.3f} ms/step "
              f"({B / res[1] * 1e3:8.0f} clips/s)")

    ce = torch.nn.CrossEntropyLoss()
    bce = torch.nn.BCEWithLogitsLoss()
    both("cfg1 notebook small-CNN LRCN 20x64x64 B=8 bf16", lambda: vc.SmallCNNLRCN(50, 20, 32, (3, 64, 64), dropout=0.5, precision="bf16"),
         torch.rand(8, 20, 3, 64, 64, device=dev), torch.randint(0, 50, (8,), device=dev), ce)
    both("cfg2 medsos ResNet-50 LRCN 16x112x112 B=8", lambda: vc.LRCN(4, 16, 32, 8, cnn_backbone="resnet50", rnn_layers=3, dropout=0.25),
         torch.rand(8, 16, 3, 112, 112, device=dev), torch.randint(0, 4, (8,), device=dev), ce)
    both("cfg3 ucf50 ResNet-50 @224 x16, 4-layer biLSTM H=56, B=2", lambda: vc.UCF50LRCN(50, 16, 56, 512, cnn_backbone="resnet50", rnn_layers=4),
         torch.rand(2, 16, 3, 224, 224, device=dev), torch.randint(0, 50, (2,), device=dev), ce)
    both("crime LRCN densenet121 full fine-tune 16x112x112 B=8",
         lambda: vc.CrimeLRCN(3, 16, 56, 512, cnn_backbone="densenet121", finetune=True, rnn_layers=4, classif_mode="multiple_binary"),
         torch.rand(8, 16, 3, 112, 112, device=dev), (torch.rand(8, 3, device=dev) > 0.5).float(), bce, steps=6)
    both("crime LRCN resnet18 full fine-tune 16x112x112 B=8",
         lambda: vc.CrimeLRCN(3, 16, 56, 512, cnn_backbone="resnet18", finetune=True, rnn_layers=4, classif_mode="multiple_binary"),
         torch.rand(8, 16, 3, 112, 112, device=dev), (torch.rand(8, 3, device=dev) > 0.5).float(), bce, steps=6)
