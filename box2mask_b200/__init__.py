"""box2mask_b200: B200-native (sm_100a) hot path of Box2Mask behind the MinkowskiEngine operator surface.

    import box2mask_b200
    box2mask_b200.install_as_minkowski_engine()   # then `import MinkowskiEngine as ME` resolves to box2mask_b200.me
"""
import sys

__all__ = ["install_as_minkowski_engine"]


def install_as_minkowski_engine():
    """Register box2mask_b200.me under the module name the reference imports (models/detection_net.py:1)."""
    from . import me
    sys.modules["MinkowskiEngine"] = me
    sys.modules["MinkowskiEngine.utils"] = me.utils
    sys.modules["MinkowskiEngine.modules"] = me.modules
    sys.modules["MinkowskiEngine.modules.resnet_block"] = me.modules.resnet_block
    return me
