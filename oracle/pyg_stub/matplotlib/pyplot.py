def _noop(*a, **k):
    return None


figure = plot = xlabel = ylabel = title = legend = savefig = close = _noop
