"""GraphedTrainStep: one whole train step (train_eval.py:20-43) replayed from a CUDA graph must give the numbers of the eager
loop, must not train the model while it is being constructed, must draw a new dropout mask per replay, and must leave no stale
derived weight copy behind for the eager calls that follow."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _batches(n, B, T, S, classes, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [(torch.rand(B, T, 3, S, S, generator=g).to(DEV), torch.randint(0, classes, (B,), generator=g).to(DEV)) for _ in range(n)]


def _eager(model, opt, crit, batches):
    losses = []
    for x, y in batches:
        opt.zero_grad(set_to_none=True)
        loss = crit(model(x), y)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    return losses


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_graphed_step_equals_eager_steps_small_cnn(precision):
    """cfg-1 model, dropout 0: the captured step and the eager step are the same kernels on the same data -> the fp32 path
    matches to 1e-5 over four optimizer steps (atomics order only), the bf16 path to 1e-2; construction leaves the weights
    untouched (bit-equal state_dict)."""
    import video_classif_b200 as vc
    torch.manual_seed(0)
    m1 = vc.SmallCNNLRCN(5, 4, 8, (3, 32, 32), dropout=0.0, precision=precision).to(DEV).train()
    m2 = copy.deepcopy(m1)
    crit = torch.nn.CrossEntropyLoss()
    batches = _batches(4, 4, 4, 32, 5)
    # SGD with momentum: an update proportional to the gradient, so equal gradients <=> equal weights (Adam's m / sqrt(v)
    # turns the rounding-noise gradients of BatchNorm-cancelled biases into +-lr steps of random sign in any implementation)
    o1 = torch.optim.SGD(m1.parameters(), lr=1e-2, momentum=0.9)
    o2 = torch.optim.SGD(m2.parameters(), lr=1e-2, momentum=0.9)
    before = {k: v.clone() for k, v in m2.state_dict().items()}
    step = vc.GraphedTrainStep(m2, o2, crit, *batches[0])
    for k, v in m2.state_dict().items():
        assert torch.equal(v, before[k]), k                       # the warm-up steps were undone
    ref = _eager(m1, o1, crit, batches)
    got = [step(x, y).item() for x, y in batches]
    tol = 1e-5 if precision == "fp32" else 1e-2
    for a, b in zip(got, ref):
        assert abs(a - b) < tol * max(1.0, abs(b)), (got, ref)
    sd1, sd2 = m1.state_dict(), m2.state_dict()
    for k in sd1:
        if sd1[k].dtype.is_floating_point:
            floor = 1e-3 if precision == "fp32" else 1e-2        # (BatchNorm biases start at 0: judged on an absolute scale)
            d = (sd1[k] - sd2[k]).abs().max().item() / max(sd1[k].abs().max().item(), floor)
            assert d < (1e-4 if precision == "fp32" else 5e-2), (k, d)
        else:
            assert torch.equal(sd1[k], sd2[k]), k                 # num_batches_tracked advanced by the replays
    with pytest.raises(ValueError):
        step(batches[0][0][:2], batches[0][1][:2])


def test_graphed_step_dropout_masks_change_and_eval_sees_current_weights():
    """Dropout 0.5: two replays on the SAME batch with a zero learning rate must give different losses (new mask per replay);
    after training replays an eager eval forward must use the CURRENT weights (no stale bf16 copy from before the replays)."""
    import video_classif_b200 as vc
    torch.manual_seed(1)
    m = vc.SmallCNNLRCN(5, 4, 8, (3, 32, 32), dropout=0.5, precision="bf16").to(DEV).train()
    crit = torch.nn.CrossEntropyLoss()
    (x, y), = _batches(1, 4, 4, 32, 5, seed=3)
    opt = torch.optim.Adam(m.parameters(), lr=0.0, capturable=True)
    step = vc.GraphedTrainStep(m, opt, crit, x, y)
    l = [step(x, y).item() for _ in range(4)]
    assert len({round(v, 6) for v in l}) > 1, l
    opt2 = torch.optim.Adam(m.parameters(), lr=5e-2, capturable=True)
    step2 = vc.GraphedTrainStep(m, opt2, crit, x, y)
    m.eval()
    with torch.no_grad():
        e0 = m(x).clone()
    m.train()
    for _ in range(3):
        step2(x, y)
    m.eval()
    with torch.no_grad():
        e1 = m(x).clone()
        m2 = copy.deepcopy(m)                                      # fresh module, same (current) weights, empty caches
        e2 = m2(x)
    assert (e1 - e0).abs().max().item() > 1e-3                    # the weights moved ...
    assert torch.allclose(e1, e2, atol=1e-5, rtol=1e-5)           # ... and the eager forward after the replays saw them


def test_graphed_step_frozen_backbone_lrcn_and_optimizer_check():
    """medsos LRCN (frozen ResNet-18, train-mode BN, dropout 0): the whole step incl. the encoder pass in one graph vs eager."""
    import video_classif_b200 as vc
    torch.manual_seed(2)
    m1 = vc.LRCN(4, 3, 16, 8, cnn_backbone="resnet18", rnn_layers=2, dropout=0.0).to(DEV).train()
    m2 = copy.deepcopy(m1)
    crit = torch.nn.CrossEntropyLoss()
    batches = _batches(3, 4, 3, 64, 4, seed=5)
    p1 = [p for p in m1.parameters() if p.requires_grad]
    p2 = [p for p in m2.parameters() if p.requires_grad]
    with pytest.raises(ValueError):
        vc.GraphedTrainStep(m2, torch.optim.Adam(p2, lr=1e-3), crit, *batches[0])          # not capturable
    o1 = torch.optim.Adam(p1, lr=1e-3)
    o2 = torch.optim.Adam(p2, lr=1e-3, capturable=True)
    step = vc.GraphedTrainStep(m2, o2, crit, *batches[0])
    ref = _eager(m1, o1, crit, batches)
    got = [step(x, y).item() for x, y in batches]
    for a, b in zip(got, ref):
        assert abs(a - b) < 2e-2 * max(1.0, abs(b)), (got, ref)
    assert torch.equal(m1.cnn_backbone.bn1.num_batches_tracked, m2.cnn_backbone.bn1.num_batches_tracked)


def test_ingest_straight_into_the_encoder_graph_input():
    """model.encoder_input_buffer(): the static input of the encoder's CUDA graph; ingest_batch(..., out=buf) followed by
    encode_async(buf) must give the features of the ordinary path (separate ingest result copied into the graph)."""
    import video_classif_b200 as vc
    from video_classif_b200.ingest import ingest_batch
    torch.manual_seed(4)
    m = vc.LRCN(4, 3, 16, 8, cnn_backbone="resnet18", rnn_layers=1, dropout=0.0).to(DEV).eval()
    m.enable_encoder_graph()
    g = torch.Generator().manual_seed(9)
    u8 = [torch.randint(0, 256, (2, 3, 64, 64, 3), generator=g, dtype=torch.uint8).to(DEV) for _ in range(3)]
    shape = (2, 3, 3, 64, 64)
    assert m.encoder_input_buffer(shape) is None                       # nothing captured yet
    with torch.no_grad():
        ref = [m(ingest_batch(u, 64, 64)).clone() for u in u8]         # ordinary path (captures the graph on first use)
        buf = m.encoder_input_buffer(shape)
        assert buf is not None and tuple(buf.shape) == shape
        for u, r in zip(u8, ref):
            side = m.side_stream(DEV)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                x = ingest_batch(u, 64, 64, out=buf)
            assert x.data_ptr() == buf.data_ptr()
            h = m.encode_async(x)
            out = m(x, features=h)
            assert torch.allclose(out, r, atol=1e-4, rtol=1e-4)
    with pytest.raises(ValueError):
        ingest_batch(u8[0], 64, 64, out=torch.empty(5, device=DEV))


def test_graphed_step_trainable_backbone_crime_lrcn():
    """crime LRCN with the WHOLE ResNet-18 trainable (lrcn.py:181-305, CONF_FINETUNE = True): forward, the backbone's backward
    kernels (weight / data gradients, BatchNorm backward, the stem) and the optimizer step replayed from one graph vs the eager
    loop, three steps with SGD (update proportional to the gradient); BatchNorm running statistics advance identically."""
    import video_classif_b200 as vc
    torch.manual_seed(6)
    m1 = vc.CrimeLRCN(3, 2, 8, 16, cnn_backbone="resnet18", finetune=True, rnn_layers=1, classif_mode="multiple_binary").to(DEV).train()
    m2 = copy.deepcopy(m1)
    crit = torch.nn.BCEWithLogitsLoss()
    g = torch.Generator().manual_seed(11)
    batches = [(torch.rand(4, 2, 3, 64, 64, generator=g).to(DEV), (torch.rand(4, 3, generator=g) > 0.5).float().to(DEV)) for _ in range(3)]
    o1 = torch.optim.SGD(m1.parameters(), lr=1e-3, momentum=0.9)
    o2 = torch.optim.SGD(m2.parameters(), lr=1e-3, momentum=0.9)
    step = vc.GraphedTrainStep(m2, o2, crit, *batches[0])
    ref = _eager(m1, o1, crit, batches)
    got = [step(x, y).item() for x, y in batches]
    for a, b in zip(got, ref):
        assert abs(a - b) < 2e-2 * max(1.0, abs(b)), (got, ref)
    sd1, sd2 = m1.state_dict(), m2.state_dict()
    assert torch.equal(sd1["cnn_backbone.bn1.num_batches_tracked"], sd2["cnn_backbone.bn1.num_batches_tracked"])
    for k in ("cnn_backbone.bn1.running_mean", "cnn_backbone.layer4.1.bn2.running_var", "cnn_backbone.conv1.weight", "fc.0.weight"):
        d = (sd1[k] - sd2[k]).abs().max().item() / max(sd1[k].abs().max().item(), 1e-3)
        assert d < 5e-2, (k, d)
