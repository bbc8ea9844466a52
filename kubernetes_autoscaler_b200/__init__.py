"""Import alias: the product package lives in ``kubernetes-autoscaler_b200/`` (a name Python cannot
import directly); this stub re-points ``__path__`` there so ``import kubernetes_autoscaler_b200.x``
resolves modules from the hyphenated directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "kubernetes-autoscaler_b200")
__path__.insert(0, _real)
