"""CPU/PyTorch restatement of the reference's `PerceptualLoss` (CycleGAN/models.py:397-476) — TEST INFRASTRUCTURE ONLY.

Only tests/ may import this module, and only as the checker of `ist_b200.model.perceptual.PerceptualLoss` (SURVEY 8f #4:
the Gram / feature kernels of the IST path re-used as a frozen-VGG training loss with batch > 1).

What is restated:
  * the VGG16 `features` stack (torchvision layout: conv/ReLU pairs in blocks of 2,2,3,3,3, each block closed by
    MaxPool2d(2,2)), walked layer by layer as models.py:432-461 does, a ReLU output being named "<block>,<index>"
    (`pool_cnt,relu_cnt`, both starting at 1; models.py:433-434,451-452);
  * the feature lists are collected in NETWORK order whatever the order of `style_layer` / `content_layer`
    (models.py:453-457), and paired with the weights by position (models.py:423-427);
  * Gram matrix G = F F^T / (h*w) per batch element (models.py:463-468);
  * loss = 1e3 * sum_i w_s[i] * MSE(G_pred[i], G_style[i]) + sum_i w_c[i] * MSE(F_pred[i], F_content[i])
    with nn.MSELoss's mean over every element including the batch (models.py:423-429).
The reference takes its weights from `torchvision.models.vgg16(pretrained=True)` (models.py:399), a download that is not
possible here; parity runs on seeded synthetic weights (oracle/synth.py) supplied through `state`.

Pinning: tests/golden/perceptual.npz holds the outputs of the UNMODIFIED reference class on seeded inputs
(oracle/make_perceptual_golden.py, which only replaces the weight download by the same synthetic weights);
tests/test_perceptual_oracle.py checks this file against them on every CPU run.
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import synth

VGG16_BLOCKS = [2, 2, 3, 3, 3]


def vgg16_conv_names():
    return ['conv%d_%d' % (b, i) for b, k in enumerate(VGG16_BLOCKS, 1) for i in range(1, k + 1)]


def get_features(state, image, content_layer, style_layer):
    """models.py:431-461: (style features, content features), each in network order."""
    output_list, content_out = [], []
    cur = image
    for b, k in enumerate(VGG16_BLOCKS, 1):
        for i in range(1, k + 1):
            name = 'conv%d_%d' % (b, i)
            cur = F.relu(F.conv2d(cur, state[name + '.weight'], state[name + '.bias'], padding=1))
            layer_name = '%d,%d' % (b, i)
            if layer_name in style_layer:
                output_list.append(cur)
            if layer_name in content_layer:
                content_out.append(cur)
        cur = F.max_pool2d(cur, kernel_size=2, stride=2)
    return output_list, content_out


def gram_matrix(feature):
    """models.py:463-468."""
    b, c, h, w = feature.size()
    Fm = feature.view(b, c, h * w)
    G = torch.bmm(Fm, Fm.transpose(1, 2))
    G.div_(h * w)
    return G


def calculate_loss(state, pred, content, style, content_layer, style_layer, weight_style, weight_content):
    """models.py:412-429."""
    pred_feature, pred_content = get_features(state, pred, content_layer, style_layer)
    _, content_target = get_features(state, content, content_layer, style_layer)
    style_target, _ = get_features(state, style, content_layer, style_layer)
    pred_gram = [gram_matrix(f) for f in pred_feature]
    target_gram = [gram_matrix(f) for f in style_target]
    style_loss = 0
    content_loss = 0
    for i in range(len(weight_style)):
        style_loss = style_loss + F.mse_loss(pred_gram[i], target_gram[i]) * weight_style[i]
    for i in range(len(weight_content)):
        content_loss = content_loss + F.mse_loss(pred_content[i], content_target[i]) * weight_content[i]
    return 1e3 * style_loss + content_loss


def loss_and_grad(state, pred, content, style, content_layer, style_layer, weight_style, weight_content):
    """Loss value and d loss / d pred (what the generator's backward receives)."""
    p = pred.detach().clone().requires_grad_(True)
    loss = calculate_loss(state, p, content, style, content_layer, style_layer, weight_style, weight_content)
    (g,) = torch.autograd.grad(loss, p)
    return loss.detach(), g


# ---- seeded cases shared by oracle/make_perceptual_golden.py and the tests ------------------------------------------
CASES = {
    # tag: (batch, H, W, style_layer, content_layer, weight_style, weight_content)
    "a": (2, 32, 48, ['1,2', '2,2', '3,3', '4,3'], ['3,3'], [1.0, 0.5, 0.25, 0.125], [2.0]),
    # names deliberately out of network order: the reference still collects features in network order (models.py:453-457)
    "b": (3, 32, 32, ['3,1', '1,1', '2,1'], ['4,2', '2,2'], [0.3, 0.2, 0.1], [1.0, 0.5]),
}


def vgg16_state(seed=0):
    """VGG16 = the first 2,2,3,3,3 convs of the VGG19 table; same per-layer streams as synth.vgg_state_dict."""
    full = synth.vgg_state_dict(seed, bias_std=0.05)
    return {k: v for k, v in full.items() if k.rsplit('.', 1)[0] in vgg16_conv_names()}


def images(b, h, w):
    out = []
    for j, base in enumerate((100, 200, 300)):
        frames = [synth.smooth_frame(max(h, w), base + i, h=h, w=w).astype(np.float32) / 127.5 - 1.0 for i in range(b)]
        out.append(np.stack([f.transpose(2, 0, 1) for f in frames]).astype(np.float32))
    return out    # pred, content, style in [-1, 1] like CycleGAN's normalised images
