"""Subset of the reference's utils.py that the model path uses (reference utils.py:42-73, :76-97)."""
from typing import Any, Callable, Dict, Optional, Tuple

from torch.nn import Module, Conv2d, ConvTranspose2d


def get_factory_adder() -> Tuple[Callable, Dict[str, Any]]:
    """Registry decorator + its dict (reference utils.py:42-73).  ``@add_class`` registers under the class
    name, ``@add_class('Name')`` under an explicit one."""
    registry: Dict[str, Any] = {}

    def add_class(obj: Any, name: Optional[str] = None) -> Any:
        if callable(obj):
            registry[obj.__name__ if name is None else name] = obj
            return obj
        return lambda cls: add_class(cls, obj)          # called with the name first

    return add_class, registry


def count_parameters(module: Module) -> int:
    return sum(p.numel() for p in module.parameters() if p.requires_grad)


def count_conv2d(module: Module) -> int:
    return sum(isinstance(m, (Conv2d, ConvTranspose2d)) for m in module.modules())


def print_net_info(net: Module) -> None:
    print('=====  Net info  =====')
    print('Layers:', count_conv2d(net))
    print('Parameters:', count_parameters(net))
    print('======================')
