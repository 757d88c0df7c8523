"""Config keys read by the hot path, with the reference's defaults.

Mirrors ``fewx.config.get_cfg()`` (fewx/config/config.py:4-100,
fewx/config/defaults.py:1-85) layered over the vendored detectron2 defaults
(d2!/config/defaults.py; fork edits: SCORE_THRESH_TEST 0.0 at :272, single cascade
stage at :321-325).  Only keys that the path (or the shipped yaml files) touch are
declared; the reference's own yaml files merge cleanly into this node.
"""
from __future__ import annotations

from ..compat import CfgNode as CN


def get_cfg() -> CN:
    _C = CN()
    _C.VERSION = 2
    _C.OUTPUT_DIR = "./output"
    _C.SEED = -1
    _C.CUDNN_BENCHMARK = False
    _C.VIS_PERIOD = 0
    _C.DEBUG = False
    _C.SAVE_DEBUG = False
    _C.SAVE_PTH = False
    _C.VIS_THRESH = 0.3
    _C.DEBUG_SHOW_NAME = False

    _C.MODEL = CN()
    _C.MODEL.DEVICE = "cuda"
    _C.MODEL.META_ARCHITECTURE = "CenterNet2Detector"
    _C.MODEL.WEIGHTS = ""
    _C.MODEL.PIXEL_MEAN = [103.530, 116.280, 123.675]
    _C.MODEL.PIXEL_STD = [1.0, 1.0, 1.0]
    _C.MODEL.MASK_ON = False
    _C.MODEL.KEYPOINT_ON = False
    _C.MODEL.LOAD_PROPOSALS = False

    _C.MODEL.BACKBONE = CN()
    _C.MODEL.BACKBONE.NAME = "build_resnet_backbone"      # detectron2's default; finetune_vovnet.yaml selects the VoVNet
    _C.MODEL.BACKBONE.FREEZE_AT = 2

    _C.MODEL.VOVNET = CN()
    _C.MODEL.VOVNET.CONV_BODY = "V-19-slim-eSE"
    _C.MODEL.VOVNET.OUT_FEATURES = ["stage2", "stage3", "stage4", "stage5"]
    _C.MODEL.VOVNET.NORM = "FrozenBN"
    _C.MODEL.VOVNET.OUT_CHANNELS = 256
    _C.MODEL.VOVNET.BACKBONE_OUT_CHANNELS = 256
    _C.MODEL.VOVNET.STAGE_WITH_DCN = (False, False, False, False)
    _C.MODEL.VOVNET.WITH_MODULATED_DCN = False
    _C.MODEL.VOVNET.DEFORMABLE_GROUPS = 1

    _C.MODEL.FPN = CN()
    _C.MODEL.FPN.IN_FEATURES = []
    _C.MODEL.FPN.OUT_CHANNELS = 256
    _C.MODEL.FPN.NORM = ""
    _C.MODEL.FPN.FUSE_TYPE = "sum"

    _C.MODEL.FCOS = CN()
    _C.MODEL.FCOS.TOP_LEVELS = 0

    _C.MODEL.PROPOSAL_GENERATOR = CN()
    _C.MODEL.PROPOSAL_GENERATOR.NAME = "RPN"
    _C.MODEL.PROPOSAL_GENERATOR.MIN_SIZE = 0

    _C.MODEL.RPN = CN()
    _C.MODEL.RPN.HEAD_NAME = "StandardRPNHead"
    _C.MODEL.RPN.IN_FEATURES = ["res4"]
    _C.MODEL.RPN.BBOX_REG_WEIGHTS = (1.0, 1.0, 1.0, 1.0)
    _C.MODEL.RPN.PRE_NMS_TOPK_TRAIN = 12000
    _C.MODEL.RPN.POST_NMS_TOPK_TRAIN = 2000
    _C.MODEL.RPN.PRE_NMS_TOPK_TEST = 1000
    _C.MODEL.RPN.POST_NMS_TOPK_TEST = 1000
    _C.MODEL.RPN.NMS_THRESH = 0.7

    _C.MODEL.ANCHOR_GENERATOR = CN()
    _C.MODEL.ANCHOR_GENERATOR.NAME = "DefaultAnchorGenerator"
    _C.MODEL.ANCHOR_GENERATOR.SIZES = [[32, 64, 128, 256, 512]]
    _C.MODEL.ANCHOR_GENERATOR.ASPECT_RATIOS = [[0.5, 1.0, 2.0]]
    _C.MODEL.ANCHOR_GENERATOR.ANGLES = [[-90, 0, 90]]
    _C.MODEL.ANCHOR_GENERATOR.OFFSET = 0.0

    _C.MODEL.RESNETS = CN()
    _C.MODEL.RESNETS.DEPTH = 50
    _C.MODEL.RESNETS.OUT_FEATURES = ["res4"]
    _C.MODEL.RESNETS.NUM_GROUPS = 1
    _C.MODEL.RESNETS.NORM = "FrozenBN"
    _C.MODEL.RESNETS.WIDTH_PER_GROUP = 64
    _C.MODEL.RESNETS.STRIDE_IN_1X1 = True
    _C.MODEL.RESNETS.RES5_DILATION = 1
    _C.MODEL.RESNETS.RES2_OUT_CHANNELS = 256
    _C.MODEL.RESNETS.STEM_OUT_CHANNELS = 64
    _C.MODEL.RESNETS.DEFORM_ON_PER_STAGE = [False, False, False, False]
    _C.MODEL.RESNETS.DEFORM_MODULATED = False
    _C.MODEL.RESNETS.DEFORM_NUM_GROUPS = 1

    c = _C.MODEL.CENTERNET = CN()
    c.NUM_CLASSES = 1
    c.IN_FEATURES = ["p3", "p4", "p5", "p6", "p7"]
    c.FPN_STRIDES = [8, 16, 32, 64, 128]
    c.PRIOR_PROB = 0.01
    c.INFERENCE_TH = 0.05
    c.CENTER_NMS = False
    c.NMS_TH_TRAIN = 0.6
    c.NMS_TH_TEST = 0.6
    c.PRE_NMS_TOPK_TRAIN = 1000
    c.POST_NMS_TOPK_TRAIN = 100
    c.PRE_NMS_TOPK_TEST = 1000
    c.POST_NMS_TOPK_TEST = 100
    c.NORM = "GN"
    c.USE_DEFORMABLE = False
    c.NUM_CLS_CONVS = 4
    c.NUM_BOX_CONVS = 4
    c.NUM_SHARE_CONVS = 0
    c.LOC_LOSS_TYPE = "giou"
    c.SIGMOID_CLAMP = 1e-4
    c.HM_MIN_OVERLAP = 0.8
    c.MIN_RADIUS = 4
    c.SOI = [[0, 80], [64, 160], [128, 320]]
    c.POS_WEIGHT = 1.0
    c.NEG_WEIGHT = 1.0
    c.REG_WEIGHT = 2.0
    c.HM_FOCAL_BETA = 4
    c.HM_FOCAL_ALPHA = 0.25
    c.LOSS_GAMMA = 2.0
    c.WITH_AGN_HM = False
    c.ONLY_PROPOSAL = False
    c.AS_PROPOSAL = False
    c.IGNORE_HIGH_FP = -1.0
    c.MORE_POS = False
    c.MORE_POS_THRESH = 0.2
    c.MORE_POS_TOPK = 9
    c.NOT_NORM_REG = True
    c.NOT_NMS = False
    c.NO_REDUCE = False

    r = _C.MODEL.ROI_HEADS = CN()
    r.NAME = "Res5ROIHeads"
    r.NUM_CLASSES = 80
    r.IN_FEATURES = ["res4"]
    r.IOU_THRESHOLDS = [0.5]
    r.IOU_LABELS = [0, 1]
    r.BATCH_SIZE_PER_IMAGE = 512
    r.POSITIVE_FRACTION = 0.25
    r.SCORE_THRESH_TEST = 0.0
    r.NMS_THRESH_TEST = 0.5
    r.PROPOSAL_APPEND_GT = True

    h = _C.MODEL.ROI_BOX_HEAD = CN()
    h.NAME = ""
    h.BBOX_REG_LOSS_TYPE = "smooth_l1"
    h.BBOX_REG_LOSS_WEIGHT = 1.0
    h.BBOX_REG_WEIGHTS = (10.0, 10.0, 5.0, 5.0)
    h.SMOOTH_L1_BETA = 0.0
    h.POOLER_RESOLUTION = 14
    h.POOLER_RESOLUTION2 = 4
    h.POOLER_SAMPLING_RATIO = 0
    h.POOLER_TYPE = "ROIAlignV2"
    h.NUM_FC = 0
    h.FC_DIM = 1024
    h.NUM_CONV = 0
    h.CONV_DIM = 256
    h.NORM = ""
    h.CLS_AGNOSTIC_BBOX_REG = True
    h.TRAIN_ON_PRED_BOXES = False
    h.MULT_PROPOSAL_SCORE = False
    h.USE_SIGMOID_CE = False
    h.PRIOR_PROB = 0.01
    h.USE_EQL_LOSS = False
    h.USE_FED_LOSS = False

    k = _C.MODEL.ROI_BOX_CASCADE_HEAD = CN()
    k.BBOX_REG_WEIGHTS = ((10.0, 10.0, 5.0, 5.0),)
    k.IOUS = (0.5,)

    _C.INPUT = CN()
    _C.INPUT.FORMAT = "BGR"
    _C.INPUT.MIN_SIZE_TRAIN = (800,)
    _C.INPUT.MAX_SIZE_TRAIN = 1333
    _C.INPUT.MIN_SIZE_TEST = 800
    _C.INPUT.MAX_SIZE_TEST = 1333
    _C.INPUT.NOT_CLAMP_BOX = False
    _C.INPUT.FS = CN()
    _C.INPUT.FS.FEW_SHOT = False
    _C.INPUT.FS.SUPPORT_WAY = 2
    _C.INPUT.FS.SUPPORT_SHOT = 10

    _C.DATASETS = CN()
    _C.DATASETS.TRAIN = ()
    _C.DATASETS.TEST = ()
    _C.DATALOADER = CN()
    _C.DATALOADER.NUM_WORKERS = 4

    s = _C.SOLVER = CN()
    s.IMS_PER_BATCH = 16
    s.BASE_LR = 0.001
    s.STEPS = (30000,)
    s.MAX_ITER = 40000
    s.WARMUP_ITERS = 1000
    s.WARMUP_FACTOR = 0.001
    s.CHECKPOINT_PERIOD = 5000
    s.HEAD_LR_FACTOR = 1.0
    s.CLIP_GRADIENTS = CN()
    s.CLIP_GRADIENTS.ENABLED = False

    _C.TEST = CN()
    _C.TEST.DETECTIONS_PER_IMAGE = 100
    _C.TEST.EVAL_PERIOD = 0
    return _C
