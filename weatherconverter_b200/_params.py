"""Parameter containers that reproduce the reference modules' dotted state_dict names without re-implementing
their nn.Module trees: a spec {dotted name: (shape, kind)} is materialised as nested anonymous modules."""
import torch
import torch.nn as nn


class Node(nn.Module):
    """Anonymous container node (children named like the reference's sub-modules, e.g. 'layer1', '0', 'bn1')."""


def register_dotted(root: nn.Module, dotted: str, tensor: torch.Tensor, buffer: bool = False):
    node = root
    parts = dotted.split(".")
    for p in parts[:-1]:
        if p not in node._modules:
            node.add_module(p, Node())
        node = node._modules[p]
    if buffer:
        node.register_buffer(parts[-1], tensor)
    else:
        node.register_parameter(parts[-1], nn.Parameter(tensor))
