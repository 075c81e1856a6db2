class SMPL: pass
from . import lbs
