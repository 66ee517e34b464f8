"""`PerceptualLoss` with the contract of the reference's CycleGAN/models.py:397-476, computed by the B200 kernels of the
IST path (the same conv / data-gradient / Gram kernels re-used as a frozen-VGG training loss, batch > 1).

Same constructor `(content_layer, style_layer, device, weight_style, weight_content)` and the same
`calculate_loss(pred, content, style) -> 0-d tensor`, differentiable w.r.t. `pred` (the generator output):
  * features are ReLU outputs of the VGG16 `features` stack named "<block>,<index>" (models.py:433-452), collected in
    NETWORK order whatever the order of the name lists (models.py:453-457) and paired with the weights by position;
  * Gram G = F F^T / (h*w) per batch element (models.py:463-468), nn.MSELoss over every element including the batch;
  * result = 1e3 * sum_i w_s[i] MSE(G_pred[i], G_style[i]) + sum_i w_c[i] MSE(F_pred[i], F_content[i]) (models.py:423-429).
The VGG16 forward, the data-gradient backward and the Gram forward / backward run in libist_b200.so (`VGG`, `GramMatrix`,
`ist_op_mse`); the network is frozen (models.py:401-402), so no weight-gradient exists.

Differences a maintainer should know:
  * weights: the reference downloads `torchvision.models.vgg16(pretrained=True)`; pass `state_dict=` (torchvision
    `features.<n>.weight` keys, bare `<n>.weight` keys or `conv{b}_{i}.weight` keys). Without it the constructor tries the
    same torchvision call and raises if the weights cannot be obtained (no silent random initialisation);
  * `content` and `style` are treated as constants (their features are computed under no_grad before `pred`'s forward; the
    plan holds one set of activations). The reference would also back-propagate into them if they required grad.
"""
import torch
import torch.nn as nn

from .. import _lib
from ..config.cfgnode import CfgNode as CN
from .meta_arch import VGG, GramMatrix

VGG16_BLOCKS = [2, 2, 3, 3, 3]
_WIDTH = [64, 128, 256, 512, 512]


def vgg16_cfg():
    """cfg.MODEL.VGG tree (same schema as IST/config/defaults.py:22-61) describing torchvision's VGG16 `features`."""
    convs, cin = {}, 3
    for b, k in enumerate(VGG16_BLOCKS, 1):
        for i in range(1, k + 1):
            convs['conv%d_%d' % (b, i)] = {'in_channels': cin, 'out_channels': _WIDTH[b - 1], 'kernel': 3, 'padding': 1}
            cin = _WIDTH[b - 1]
    cfg = CN()
    cfg.MODEL = CN()
    cfg.MODEL.VGG = CN()
    cfg.MODEL.VGG.CONV_LAYERS_DICT = [convs]
    cfg.MODEL.VGG.POOL_LAYERS_DICT = [{'pool_%d' % b: {'kernel_size': 2, 'stride': 2} for b in range(1, 6)}]
    cfg.MODEL.VGG.FORWARD_SEQ = [n for b, k in enumerate(VGG16_BLOCKS, 1)
                                 for n in ['conv%d_%d' % (b, i) for i in range(1, k + 1)] + ['pool_%d' % b]]
    cfg.MODEL.VGG.OUT_SEQ = [n for b, k in enumerate(VGG16_BLOCKS, 1)
                             for n in ['relu%d_%d' % (b, i) for i in range(1, k + 1)] + ['pool_%d' % b]]
    return cfg


def _features_index_to_name():
    """torchvision vgg16.features index of every conv -> conv{b}_{i}."""
    out, idx = {}, 0
    for b, k in enumerate(VGG16_BLOCKS, 1):
        for i in range(1, k + 1):
            out[idx] = 'conv%d_%d' % (b, i)
            idx += 2                      # conv, ReLU
        idx += 1                          # pool
    return out


def convert_state_dict(state):
    """Accept torchvision (`features.0.weight` / `0.weight`) or IST-style (`conv1_1.weight`) keys; returns IST-style keys."""
    names = _features_index_to_name()
    out = {}
    for k, v in state.items():
        parts = k.split('.')
        if parts[0] == 'classifier':
            continue
        if parts[0] == 'features':
            parts = parts[1:]
        if parts[0].isdigit():
            n = int(parts[0])
            if n not in names:
                raise KeyError(k)
            parts = [names[n]] + parts[1:]
        out['.'.join(parts)] = v
    return out


class _BatchMSE(torch.autograd.Function):
    """nn.MSELoss()(x, t) for feature maps [b,C,H,W] (mean over every element) through `ist_op_mse`."""

    @staticmethod
    def forward(ctx, x, t):
        x = x.contiguous().float()
        t = t.detach().contiguous().float()
        b, c, h, w = x.shape
        per_frame = torch.empty(b, 2, device=x.device, dtype=torch.float32)
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _lib.check(_lib.load().ist_op_mse(_lib.ptr(x), _lib.ptr(t), 1.0 / b, _lib.ptr(per_frame), _lib.ptr(dx), b, c, h, w,
                                              _lib.stream_ptr(x.device)))
        ctx.save_for_backward(dx)
        return per_frame[:, 0].sum()

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return dx * g, None


class PerceptualLoss():
    def __init__(self, content_layer, style_layer, device, weight_style, weight_content, state_dict=None):
        device = torch.device(device)
        if device.type != 'cuda':
            raise _lib.IstError("PerceptualLoss runs on CUDA devices only: this package has no CPU path")
        if state_dict is None:
            try:
                import torchvision.models as models
                state_dict = models.vgg16(pretrained=True).features.state_dict()      # models.py:399
            except Exception as e:
                raise _lib.IstError("VGG16 weights unavailable (%s); pass state_dict= with torchvision vgg16 weights" % (e,))
        self.net = VGG(vgg16_cfg(), 'max')
        self.net.load_state_dict(convert_state_dict(state_dict))
        self.net.to(device)
        for param in self.net.parameters():
            param.requires_grad = False
        self.device = device
        self.style_layer = style_layer
        self.content_layer = content_layer
        self.weight_style = weight_style
        self.weight_content = weight_content
        self.content_loss_func = [nn.MSELoss()] * len(content_layer)
        self.style_loss_func = [nn.MSELoss()] * len(style_layer)
        self._gram = GramMatrix()

    # "<block>,<index>" -> relu{block}_{index}, in network order, restricted to names that exist in VGG16
    def _keys(self, layer_names):
        wanted = set(layer_names)
        keys = []
        for b, k in enumerate(VGG16_BLOCKS, 1):
            for i in range(1, k + 1):
                if '%d,%d' % (b, i) in wanted:
                    keys.append('relu%d_%d' % (b, i))
        return keys

    def _get_features(self, image, content_layer, style_layer):
        """models.py:431-461: (style features, content features), each in network order; one VGG forward."""
        sk, ck = self._keys(style_layer), self._keys(content_layer)
        keys = list(dict.fromkeys(sk + ck))
        if not keys:
            return [], []
        feats = dict(zip(keys, self.net(image, keys)))
        return [feats[k] for k in sk], [feats[k] for k in ck]

    def calculate_loss(self, pred, content, style):
        with torch.no_grad():
            _, content_target = self._get_features(content, self.content_layer, self.style_layer)
            style_target, _ = self._get_features(style, self.content_layer, self.style_layer)
            target_gram = [self._gram(f) for f in style_target]
        pred_feature, pred_content = self._get_features(pred, self.content_layer, self.style_layer)
        pred_gram = [self._gram(f) for f in pred_feature]

        style_loss = 0
        content_loss = 0
        for i in range(len(self.weight_style)):
            style_loss += self.style_loss_func[i](pred_gram[i], target_gram[i]) * self.weight_style[i]
        for i in range(len(self.weight_content)):
            content_loss += _BatchMSE.apply(pred_content[i], content_target[i]) * self.weight_content[i]
        return 1e3 * style_loss + content_loss
