"""Drop-in package name: ``import RVGP`` resolves to the B200-native implementation in ``rvgp_b200``.

Same public names as the reference's RVGP/__init__.py:1-3."""
from rvgp_b200 import kernels  # noqa: F401
from rvgp_b200.dataclass import data as create_data_object  # noqa: F401
from rvgp_b200.main import train_gp as fit  # noqa: F401
