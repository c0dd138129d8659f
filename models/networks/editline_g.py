"""DeepFillC2Generator = netG (reference models/networks/editline_g.py:13-221): coarse encoder + style
encoder with global pooling, coarse decoder, then a hallucination branch and a patch-match branch with
contextual attention feeding a joint decoder. Same option flags and state_dict keys as the reference;
``forward`` is one C-ABI call (``se_netG_forward``)."""
import warnings

from models.networks.base_network import BaseNetwork
from models.networks.editline2_g import _build_layers
from models.networks.splitcam import ReduceContextAttentionP1, ReduceContextAttentionP2


class DeepFillC2Generator(BaseNetwork):
    NET_ID = "G"

    @staticmethod
    def modify_commandline_options(parser, is_train):
        parser.add_argument("--use_cam", action="store_true", help="use the contextual attention module")
        parser.add_argument("--pool_type", default="avg", help="global pooling of the style encoder: avg | max")
        parser.add_argument("--no_mask_cc", action="store_true", help="do not mask the style-encoder input")
        parser.add_argument("--no_mask_coarse", action="store_true", help="do not blend the coarse result with the input")
        return parser

    def __init__(self, opt):
        super().__init__()
        self.opt = opt
        self.no_mask_coarse = opt.no_mask_coarse
        self.pool_type = opt.pool_type
        self.use_cam = opt.use_cam
        self.no_mask_cc = opt.no_mask_cc
        self.cnum = 48
        self.precision = getattr(opt, "precision", "bf16")
        self.cam_1 = ReduceContextAttentionP1(nn_hard=False, ufstride=2, stride=2, bkg_patch_size=4, pd=0,
                                              is_th=True, th=0.1, norm_type=1)
        self.cam_2 = ReduceContextAttentionP2(ufstride=2, bkg_patch_size=4, stride=2, pd=0, mk=False)
        _build_layers(self, "G")

    def _configure_engine(self, eng):
        eng.set_options(use_cam=self.use_cam, pool_type=self.pool_type, no_mask_cc=self.no_mask_cc,
                        no_mask_coarse=self.no_mask_coarse, joint_train_inp=self.opt.joint_train_inp)

    def get_param_list(self, stage="all"):
        named = list(self.named_parameters())
        if stage in ("all", "image"):
            return [p for _, p in named]
        if stage == "coarse":
            return [p for n, p in named if n.startswith("conv")]
        if stage == "fine":
            return [p for n, p in named if not n.startswith("conv")]
        warnings.warn("no generator param update")
        return []

    def forward(self, x, x2, mask, mask2, guide=None):
        # guide=None (reference :127-130: an all-ones sketch channel) is handled inside the library
        return self.engine().netG(x.float(), x2.float(), mask.float(), mask2.float(), None if guide is None else guide.float(),
                                  precision=self.precision)
