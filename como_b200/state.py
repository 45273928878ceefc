"""`WindowState`: plain attribute bag carrying the reference Mapping's state names (kf_poses, Knm_Kmminv, ...).
Lives in its own module so that building a window with CPU generators never loads the CUDA library."""


class WindowState:
    def __init__(self, **kw):
        self.__dict__.update(kw)
