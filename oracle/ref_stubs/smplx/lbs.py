def batch_rodrigues(*a, **k): raise NotImplementedError
def blend_shapes(*a, **k): raise NotImplementedError
def vertices2joints(*a, **k): raise NotImplementedError
def batch_rigid_transform(*a, **k): raise NotImplementedError
def transform_mat(*a, **k): raise NotImplementedError
