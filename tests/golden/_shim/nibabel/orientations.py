def aff2axcodes(affine):
    return ("R", "A", "S")
