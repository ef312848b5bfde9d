"""Oracle restatement of the eager steps either side of the purification call -- test infrastructure only.

SDE_Adv_Model.forward (eval_sde_adv.py:68-93): ImageNet inputs are resized 224 -> 256 with
F.interpolate(mode='bilinear', align_corners=False), every input is mapped [0,1] -> [-1,1] before the runner, the
purified image is resized back to 224 and mapped to [0,1] before the classifier; the classifier wrappers
(utils.py:144-153) then apply (x - mu) / sigma. PyTorch's own interpolate IS the reference implementation here."""
import torch
import torch.nn.functional as F

IMAGENET_MU = (0.485, 0.456, 0.406)
IMAGENET_SIGMA = (0.229, 0.224, 0.225)


def pre(x01, model_hw=None):
    """[0,1] images -> the runner's input: optional bilinear resize to the model grid, then (x - 0.5) * 2."""
    if model_hw is not None and tuple(x01.shape[2:]) != tuple(model_hw):
        x01 = F.interpolate(x01, size=tuple(model_hw), mode='bilinear', align_corners=False)
    return (x01 - 0.5) * 2


def post(x_re, out_hw=None, norm=None):
    """purified [-1,1] images -> classifier input: optional bilinear resize, (x + 1) / 2, optional (x - mu) / sigma."""
    if out_hw is not None and tuple(x_re.shape[2:]) != tuple(out_hw):
        x_re = F.interpolate(x_re, size=tuple(out_hw), mode='bilinear', align_corners=False)
    x = (x_re + 1) * 0.5
    if norm is not None:
        mu = torch.tensor(norm[0]).view(1, 3, 1, 1)
        sigma = torch.tensor(norm[1]).view(1, 3, 1, 1)
        x = (x - mu) / sigma
    return x
