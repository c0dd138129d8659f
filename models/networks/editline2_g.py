"""MDGenerator = netM, the mask predictor (reference models/networks/editline2_g.py:13-94): a shared
10-layer gated-conv encoder, an image decoder (tanh) and a mask decoder (sigmoid). Attribute names are
the reference's state_dict keys; ``forward`` is one C-ABI call (``se_netM_forward``)."""
import torch

from models.networks.base_network import BaseNetwork
from models.networks.utils import bind_layers, gen_conv, gen_deconv
from sketchedit_b200.arch import NET_LAYERS


def _build_layers(module, net_id):
    for l in NET_LAYERS[net_id]:
        if l.kind == "deconv":
            layer = gen_deconv(l.cin, l.cout)
        else:
            act = {"elu": torch.nn.ELU(), "relu": torch.nn.ReLU(), None: None}[l.act]
            if l.cout == 3:
                act = None
            layer = gen_conv(l.cin, l.cout, l.k, l.stride, rate=l.rate, activation=act)
        setattr(module, l.name, layer)
    bind_layers(module)


class MDGenerator(BaseNetwork):
    NET_ID = "M"

    def __init__(self, opt):
        super().__init__()
        self.precision = getattr(opt, "precision", "bf16")
        _build_layers(self, "M")

    def get_param_list(self, stage="all"):
        if stage in ("all", "mask"):
            return [p for _, p in self.named_parameters()]
        if stage == "maskim":
            return [p for n, p in self.named_parameters() if n.startswith("conv")]
        return []

    def forward(self, x, guide):
        mask1, x_stage1 = self.engine().netM(x.float(), guide.float(), precision=self.precision, want_image=True)
        return mask1, x_stage1
