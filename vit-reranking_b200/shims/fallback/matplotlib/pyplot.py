def __getattr__(name):
    raise RuntimeError(f"matplotlib.pyplot.{name}: matplotlib is not installed (vitrerank shim)")
