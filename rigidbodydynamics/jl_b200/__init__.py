from .spatial import *          # noqa
from .joint_types import *      # noqa
from .mechanism import *        # noqa
from .urdf import parse_urdf, default_urdf_joint_types
