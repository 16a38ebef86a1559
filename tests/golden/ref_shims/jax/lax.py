def pmean(x, axis_name=None):
    return x
