"""Host-side checks of the external-classifier handling (no GPU): BatchNorm folding keeps the function and the user's module."""
import copy

import pytest
import torch


def _randomise_bn(model, seed):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(0.1 * torch.randn(m.num_features, generator=g))
            m.running_var.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.weight.data.copy_(0.5 + torch.rand(m.num_features, generator=g))
            m.bias.data.copy_(0.1 * torch.randn(m.num_features, generator=g))


@pytest.mark.parametrize("name", ["resnet18", "inception_v3", "vgg16"])
def test_fold_batchnorm_keeps_logits_gradients_and_the_users_module(name):
    from spaa_b200.classifier import Classifier, fold_batchnorm
    clf = Classifier(name, "cpu", [0], allow_random_init=True)
    _randomise_bn(clf.model, 3)
    before = copy.deepcopy(clf.model.state_dict())
    from spaa_b200.classifier import ConvBiasAct, FusedBasicBlock, FusedReLUMaxPool2d
    if name == "vgg16":                                   # no BatchNorm and no fusion asked for: nothing to change, the classifier itself comes back
        assert fold_batchnorm(clf, fuse_pool=False, fuse_bias=False, fuse_stem=False) is clf
    view = fold_batchnorm(clf)
    assert view is not clf and view.input_sz == clf.input_sz
    assert not any(isinstance(m, torch.nn.BatchNorm2d) for m in view.model.modules())
    # the private copy's ReLU -> MaxPool2d pairs / lone MaxPool2d modules are the fused module (on CPU tensors it runs the stock ops)
    fused = [m for m in view.model.modules() if isinstance(m, FusedReLUMaxPool2d)]
    assert not any(type(m) is torch.nn.MaxPool2d for m in view.model.modules())
    assert [(m.kernel_size, m.stride, m.padding, m.with_relu) for m in fused] == {
        "resnet18": [(3, 2, 1, True)], "vgg16": [(2, 2, 0, True)] * 5, "inception_v3": [(3, 2, 0, False)] * 2}[name]
    assert any(type(m) is torch.nn.MaxPool2d for m in clf.model.modules())
    from spaa_b200.classifier import S2DStem
    assert getattr(view, "stem_s2d", False) == (name == "resnet18") and (type(getattr(view.model, "conv1", None)) is S2DStem) == (name == "resnet18")
    # every cuDNN convolution of the copy lost its bias to the fused kernel behind it (pooling kernel or ConvBiasAct)
    kinds = [type(m) for m in view.model.modules()]
    assert (kinds.count(FusedBasicBlock), kinds.count(ConvBiasAct)) == {"resnet18": (8, 16), "vgg16": (0, 8), "inception_v3": (0, 96)}[name]
    assert not any(type(m) is torch.nn.Conv2d and m.bias is not None for m in view.model.modules())
    assert fold_batchnorm(clf, fuse_pool=False, fuse_bias=False, fuse_stem=False) is not view
    assert all(fm.bias is not None for fm in fused) == (name != "inception_v3")
    for k, v in clf.model.state_dict().items():           # the user's network is untouched
        assert torch.equal(v, before[k])
    x = torch.rand(2, 3, *clf.input_sz, generator=torch.Generator().manual_seed(5))
    outs, grads = [], []
    for net in (clf.model, view.model):
        leaf = x.clone().requires_grad_(True)
        y = net(leaf)
        y = y.logits if hasattr(y, "logits") else y
        g, = torch.autograd.grad(y[:, 7].sum(), leaf)
        outs.append(y.detach()); grads.append(g)
    scale = outs[0].abs().max().item()
    assert (outs[0] - outs[1]).abs().max().item() <= 2e-5 * max(scale, 1.0)            # fp32 re-association only
    assert torch.equal(outs[0].argmax(1), outs[1].argmax(1))
    # random-init inception_v3 (94 conv layers, no trained scales) has input gradients of ~1e-11 whose rounding noise is amplified
    # layer by layer: it is held to a relative Frobenius bound, resnet18 to max-abs
    if name == "vgg16":                                   # same stock ops (the bias added after the convolution instead of inside it)
        assert (grads[0] - grads[1]).abs().max().item() <= 1e-5 * grads[0].abs().max().item()
    elif name == "resnet18":
        assert (grads[0] - grads[1]).abs().max().item() <= 1e-4 * grads[0].abs().max().item()
    else:
        rel = ((grads[0] - grads[1]).double().norm() / grads[0].double().norm()).item()
        assert rel <= 5e-2, rel


def test_fold_batchnorm_leaves_opaque_and_training_classifiers_alone():
    from spaa_b200.classifier import fold_batchnorm

    class Opaque:
        def __call__(self, im, crop):
            return im

    o = Opaque()
    assert fold_batchnorm(o) is o

    class Training:
        model = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4)).train()
        input_sz = (8, 8)

    t = Training()
    assert fold_batchnorm(t) is t


@pytest.mark.parametrize("hw", [(16, 20), (15, 17), (224, 224)])
def test_s2d_stem_equals_the_7x7_stride_2_convolution(hw):
    """The re-parameterised stem (4x4 stride-1 convolution over the 2x2 space-to-depth fold) against the convolution it replaces: values and
    input gradient, even sizes (folded) and odd sizes (the original convolution runs)."""
    from spaa_b200.classifier import S2DStem, s2d_fold
    torch.manual_seed(1)
    conv = torch.nn.Conv2d(3, 8, 7, 2, 3)
    stem = S2DStem(conv)
    x = torch.randn(2, 3, *hw)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = conv(xa), stem(xb)
    assert ya.shape == yb.shape and (ya - yb).abs().max().item() <= 1e-5
    cot = torch.randn_like(ya)
    ya.backward(cot); yb.backward(cot)
    assert (xa.grad - xb.grad).abs().max().item() <= 1e-5
    if hw[0] % 2 == 0 and hw[1] % 2 == 0:
        f = s2d_fold(x)
        assert f.shape == (2, 16, hw[0] // 2 + 3, hw[1] // 2 + 3)
        assert torch.equal(f[:, 12:], torch.zeros_like(f[:, 12:])) and torch.equal(f[:, :, :2], torch.zeros_like(f[:, :, :2]))
        assert torch.equal(f[:, 0:3, 2, 2], x[:, :, 0, 0]) and torch.equal(f[:, 9:12, 2, 2], x[:, :, 1, 1]) and torch.equal(f[:, 3:6, 3, 2], x[:, :, 2, 1])
        assert (stem(f) - ya.detach()).abs().max().item() <= 1e-5


def test_classifier_requires_pretrained_weights_unless_opted_in(tmp_path, monkeypatch):
    """The reference always loads the exact ImageNet weights (classifier.py:36); a missing checkpoint must not silently become a random network."""
    import pytest
    import torch
    from spaa_b200.classifier import Classifier
    monkeypatch.delenv("SPAA_WEIGHTS_DIR", raising=False)
    monkeypatch.delenv("SPAA_ALLOW_RANDOM_INIT", raising=False)
    monkeypatch.setattr(torch.hub, "get_dir", lambda: str(tmp_path))
    with pytest.raises(FileNotFoundError):
        Classifier("resnet18", "cpu", [0])
    with pytest.warns(UserWarning):
        c = Classifier("resnet18", "cpu", [0], allow_random_init=True)
    assert c.pretrained is False
