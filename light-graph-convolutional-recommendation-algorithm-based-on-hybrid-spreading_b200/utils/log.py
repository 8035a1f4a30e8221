"""Drop-in for /root/reference/utils/log.py: module-level `logger` writing to the console and to
cfg.LOG["file_path"]/app_<unix-ts>.log (same names and levels, reference log.py:14-97)."""
import datetime
import logging

from const import cfg


class Logger:
    def __init__(self, save_path: str = None) -> None:
        self.logger = logging.getLogger("logger")
        self.logger.setLevel(logging.DEBUG)
        if self.logger.handlers:      # re-import safe: do not stack handlers
            return
        fmt = logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s")
        console = logging.StreamHandler()
        console.setLevel(logging.DEBUG)
        console.setFormatter(fmt)
        self.logger.addHandler(console)
        if save_path is not None:
            stamp = int(datetime.datetime.now().timestamp())
            fh = logging.FileHandler(save_path + "app_" + str(stamp) + ".log", encoding="utf-8")
            fh.setLevel(logging.INFO)
            fh.setFormatter(fmt)
            self.logger.addHandler(fh)

    def debug(self, message: str):
        self.logger.debug(message)

    def info(self, message: str):
        self.logger.info(message)

    def warning(self, message: str):
        self.logger.warning(message)

    def error(self, message: str):
        self.logger.error(message)

    def critical(self, message: str):
        self.logger.critical(message)


logger = Logger(save_path=cfg.LOG["file_path"])
