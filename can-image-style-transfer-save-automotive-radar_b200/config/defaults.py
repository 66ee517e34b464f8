"""Default configuration tree: same keys and values as the reference's IST/config/defaults.py:7-104 (the config keys are
the API of the path; tests/test_config.py compares this tree with the reference's, captured in tests/golden/cfg_defaults.json)."""
from .cfgnode import CfgNode as CN

_C = CN()

# model ------------------------------------------------------------------------------------------------------------------
_C.MODEL = CN()
_C.MODEL.META_ARCHITECTURE = 'VGG'
_C.MODEL.DEVICE = 'cuda'
_C.MODEL.MODELS_DIR = './models'
_C.MODEL.WEIGHTS = './models/vgg_conv.pth'

_VGG_CHANNELS = [('conv1_1', 3, 64), ('conv1_2', 64, 64),
                 ('conv2_1', 64, 128), ('conv2_2', 128, 128),
                 ('conv3_1', 128, 256), ('conv3_2', 256, 256), ('conv3_3', 256, 256), ('conv3_4', 256, 256),
                 ('conv4_1', 256, 512), ('conv4_2', 512, 512), ('conv4_3', 512, 512), ('conv4_4', 512, 512),
                 ('conv5_1', 512, 512), ('conv5_2', 512, 512), ('conv5_3', 512, 512), ('conv5_4', 512, 512)]
_C.MODEL.VGG = CN()
_C.MODEL.VGG.CONV_LAYERS_DICT = [{
    name: {'in_channels': cin, 'out_channels': cout, 'kernel': 3, 'padding': 1} for name, cin, cout in _VGG_CHANNELS
}]
_C.MODEL.VGG.POOL_LAYERS_DICT = [{'pool_%d' % i: {'kernel_size': 2, 'stride': 2} for i in range(1, 6)}]
_BLOCKS = [2, 2, 4, 4, 4]
_C.MODEL.VGG.FORWARD_SEQ = [n for b, k in enumerate(_BLOCKS, 1)
                            for n in ['conv%d_%d' % (b, i) for i in range(1, k + 1)] + ['pool_%d' % b]]
_C.MODEL.VGG.OUT_SEQ = [n for b, k in enumerate(_BLOCKS, 1)
                        for n in ['relu%d_%d' % (b, i) for i in range(1, k + 1)] + ['pool_%d' % b]]

# loss -------------------------------------------------------------------------------------------------------------------
_C.LOSS = CN()
_C.LOSS.CONTENT_LAYERS = ['relu4_2']
_C.LOSS.STYLE_LAYERS = ['relu1_1', 'relu2_1', 'relu3_1', 'relu4_1', 'relu5_1']
_C.LOSS.CONTENT_WEIGHTS = [5e-1]
_C.LOSS.STYLE_WEIGHTS = [1e3 / n ** 2 for n in [64, 128, 256, 512, 512]]
_C.LOSS.MAX_ITER = 300
_C.LOSS.LOG_ITER_SHOW = 0.1

_C.HRLOSS = CN()
_C.HRLOSS.MAX_ITER = 500

# data -------------------------------------------------------------------------------------------------------------------
_C.DATA = CN()
_C.DATA.STYLE_IMG_PATH = '/home/dj/Downloads/lidar/nuscene/save_new/lidar/00043.png'
_C.DATA.CONTENT_IMG_PATH = '/home/dj/Downloads/lidar/nuscene/save_new/radar/00043.png'
_C.DATA.IMG_SIZE = 512
_C.DATA.IMAGENET_MEAN = [0.40760392, 0.45795686, 0.48501961]
_C.HRDATA = CN()
_C.HRDATA.IMG_SIZE = 512

# output -----------------------------------------------------------------------------------------------------------------
_C.OUTPUT = CN()
_C.OUTPUT.DIR = './output/full_transfer/'
_C.OUTPUT.FILE_NAME = 'res.jpg'
_C.OUTPUT.HR_FILE_NAME = 'hr_res.jpg'


def get_cfg_defaults():
    """A clone, so the defaults are never altered (IST/config/defaults.py:100-104)."""
    return _C.clone()
