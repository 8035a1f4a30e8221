"""Drop-in for /root/reference/utils/picture.py (plotMetric, picture.py:11-27).  Plotting is not
on the hot path; when matplotlib is absent the call is a logged no-op instead of an ImportError."""


def plotMetric(xpoints: list, ypoints: list, xlabel: str, ylabel: str, title: str, save_path: str) -> None:
    try:
        import matplotlib

        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        from utils.log import logger

        logger.info(f"matplotlib not installed - skipping plot {save_path}")
        return
    plt.figure()
    plt.plot(xpoints, ypoints)
    plt.xlabel(xlabel)
    plt.ylabel(ylabel)
    plt.title(title)
    plt.savefig(save_path)
    plt.close()
