class SpatialImage:
    pass
