"""GPU: wespeaker ResNet speaker encoder on libwesep_b200 vs the fp64 oracle restatement
(oracle/resnet.py — PARITY UNPINNED by the reference: wespeaker is an external package, see the oracle header), and pBSRNN
with `joint_training=True` end to end."""
import pytest
import torch

from oracle import bsrnn as ob
from oracle import resnet as orn
from tests.test_gpu_kernels import check, rnd

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _fill(m, seed):
    from wesep_b200 import synth
    synth.fill_state_dict_(m.state_dict(), seed=seed)
    for k, v in m.state_dict().items():          # positive running variances / BatchNorm scales around 1
        if k.endswith("running_var"):
            v.abs_().add_(0.5)
        if k.endswith("num_batches_tracked"):
            v.zero_()


@pytest.mark.parametrize("stride", [1, 2])
def test_im2col_conv3x3(stride):
    """im2col + pointwise GEMM == F.conv2d(3x3, pad 1, stride), forward and both gradients."""
    import torch.nn.functional as F
    from wesep_b200 import ops
    n, C, H, W, Co = 3, 5, 9, 14, 8
    x0 = rnd(n, C, H, W, seed=1)
    w0 = rnd(Co, C, 3, 3, seed=2, scale=0.3)
    x = ops.new_act(n, C, H * W, DEV)
    x.copy_(x0.reshape(n, C, H * W))
    x.requires_grad_(True)
    w = w0.clone().requires_grad_(True)
    col = ops.Im2Col3x3Fn.apply(x, H, W, stride)
    w2d = w.reshape(Co, -1)
    y = ops.conv1x1_bigk(col, F.pad(w2d, (0, col.shape[1] - w2d.shape[1])))
    x64, w64 = x0.double().requires_grad_(True), w0.double().requires_grad_(True)
    ref = F.conv2d(x64, w64, None, stride, 1)
    check("y", y, ref.reshape(n, Co, -1), 1e-5)
    g = rnd(*y.shape, seed=3)
    y.backward(g)
    ref.backward(g.double().reshape(ref.shape))
    check("dx", x.grad, x64.grad.reshape(n, C, H * W), 3e-5)
    check("dw", w.grad, w64.grad, 2e-4)


@pytest.mark.parametrize("blocks,m,feat,T", [((1, 1, 1, 1), 8, 16, 40), ((2, 1, 2, 1), 16, 24, 57)])
def test_resnet_vs_oracle(blocks, m, feat, T):
    from wesep_b200.modules.speaker.resnet import BasicBlock, ResNet
    net = ResNet(BasicBlock, list(blocks), m_channels=m, feat_dim=feat, embed_dim=32, pooling_func="TSTP", two_emb_layer=False)
    _fill(net, 7)
    net = net.to(DEV).train()
    sd64 = {k: v.detach().clone().double() for k, v in net.state_dict().items()}
    names = [k for k, _ in net.named_parameters()]
    P64 = {k: sd64[k].requires_grad_(True) for k in names}
    sd64.update(P64)
    x0 = rnd(4, T, feat, seed=3)
    _, emb = net(x0)
    bufs = {}
    ref = orn.resnet_forward(sd64, x0.double(), num_blocks=blocks, training=True, buffers_out=bufs)
    check("emb", emb, ref, 2e-5)
    g = rnd(*emb.shape, seed=4)
    emb.backward(g)
    ref.backward(g.double())
    tot = sum(float(P64[k].grad.norm()) ** 2 for k in names) ** 0.5
    for k, p in net.named_parameters():
        r = P64[k].grad
        if float(r.norm()) > 1e-5 * tot:
            check("grad " + k, p.grad, r, 1e-3)
    for k, v in bufs.items():                       # running statistics updated like nn.BatchNorm2d
        check("buffer " + k, net.state_dict()[k], v, 1e-5)
    assert int(net.bn1.num_batches_tracked) == 1


def test_bsrnn_joint_training_recipe_constructs_and_steps():
    """`get_model("BSRNN")(**bsrnn.yaml model_args)` (joint_training: True, ResNet34 TSTP) builds and runs a train step; the
    estimate agrees with the oracle pBSRNN fed by the oracle ResNet embedding (small network for the comparison)."""
    from oracle import losses as olosses
    from wesep_b200 import ops, synth
    from wesep_b200.models import get_model
    spk_args = dict(feat_dim=80, embed_dim=256, pooling_func="TSTP", two_emb_layer=False)
    args = dict(sr=16000, win=512, stride=128, feature_dim=16, num_repeat=1, spk_fuse_type="multiply", use_spk_transform=False,
                multi_fuse=False, joint_training=True, spk_model="ResNet34", spk_model_init=False, spk_args=spk_args,
                spk_emb_dim=256, spk_feat=True, feat_type="consistent", multi_task=True, spksInTrain=11)
    m = get_model("BSRNN")(**args)
    _fill(m, 5)
    m = m.to(DEV).train()
    b = synth.make_batch(2, T=4000, Te=8, seed=9, device=DEV)
    fb = rnd(2, 60, 80, seed=6)
    est, logits = m(b["wav_mix"], fb)
    assert est.shape == (2, 4000) and logits.shape == (2, 11)
    sd64 = {k: v.detach().double() for k, v in m.state_dict().items()}
    sd64 = {k: v.cpu() for k, v in sd64.items()}
    emb64 = orn.resnet_forward({k[len("spk_model."):]: v for k, v in sd64.items() if k.startswith("spk_model.")}, fb.double().cpu(),
                               training=True)
    ref = ob.bsrnn_forward(sd64, b["wav_mix"].double().cpu(), emb64, num_repeat=1, spk_fuse_type="multiply", multi_fuse=False)
    check("est", est, ref.to(DEV), 2e-3)
    rows = olosses.sisdr_per_row(est.detach().double(), b["wav_targets"].double())
    rows64 = olosses.sisdr_per_row(ref.to(DEV), b["wav_targets"].double())
    assert float((rows - rows64).abs().max()) <= 0.01
    check("logits", logits, (emb64 @ sd64["pred_linear.weight"].t() + sd64["pred_linear.bias"]).to(DEV), 1e-4)
    losses, _ = ops.sisdr_losses([est], b["wav_targets"])
    (losses[0] + 0.5 * ops.cross_entropy(logits, torch.tensor([1, 3], device=DEV))).backward()
    missing = [k for k, p in m.named_parameters() if p.grad is None]
    assert not missing, missing[:5]


def test_bsrnn_yaml_model_args_construct():
    """The recipe's model_args (examples/librimix/tse/v2/confs/bsrnn.yaml:46-83) construct unchanged: 28.07 M parameters."""
    from wesep_b200.models import get_model
    m = get_model("BSRNN")(sr=16000, win=512, stride=128, feature_dim=128, num_repeat=6, spk_fuse_type="multiply",
                           use_spk_transform=False, multi_fuse=False, joint_training=True, spk_model="ResNet34",
                           spk_model_init=False, spk_args=dict(feat_dim=80, embed_dim=256, pooling_func="TSTP", two_emb_layer=False),
                           spk_emb_dim=256, spk_model_freeze=False, spk_feat=True, feat_type="consistent", multi_task=False)
    assert sum(p.numel() for p in m.parameters()) == 28069608


def test_bsrnn_joint_training_raw_enrollment():
    """`spk_feat: False`, `feat_type: consistent` (bsrnn.yaml:15 alternative): the model computes the enrollment features itself
    (pre-emphasis + log-mel + mean removal, no gradient) and trains the speaker encoder through them; equals feeding the same
    features to the `spk_feat: True` model."""
    from wesep_b200 import synth
    from wesep_b200.models import get_model
    from wesep_b200.modules.speaker.consistent import consistent_features
    args = dict(sr=16000, win=512, stride=128, feature_dim=16, num_repeat=1, spk_fuse_type="multiply", use_spk_transform=False,
                multi_fuse=False, joint_training=True, spk_model="ResNet18", spk_model_init=False,
                spk_args=dict(feat_dim=80, embed_dim=256, pooling_func="TSTP", two_emb_layer=False), spk_emb_dim=256,
                spk_model_freeze=False, feat_type="consistent", multi_task=False)
    torch.manual_seed(0)
    m_raw = get_model("BSRNN")(**dict(args, spk_feat=False)).to(DEV).train()
    m_fb = get_model("BSRNN")(**dict(args, spk_feat=True)).to(DEV).train()
    sd = {k: v for k, v in m_raw.state_dict().items() if not k.startswith(("preEmphasis", "spk_encoder"))}
    m_fb.load_state_dict(sd)
    b = synth.make_batch(2, T=6000, Te=9000, seed=3, device=DEV)
    est_raw, emb_raw = m_raw(b["wav_mix"], b["spk_embeds"])
    feats = consistent_features(b["spk_embeds"], m_raw.preEmphasis, m_raw.spk_encoder)
    assert feats.shape == (2, 1 + 9000 // 128, 80)
    est_fb, emb_fb = m_fb(b["wav_mix"], feats)
    assert torch.equal(est_raw, est_fb) and torch.equal(emb_raw, emb_fb)
    est_raw.square().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m_raw.parameters())
