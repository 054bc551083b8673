"""Logger / config helpers (mirror of utils/auxiliary.py of the reference)."""
import logging
import os

import yaml


class Logger:
    def __init__(self, name="gsevt", log_file=None, level=logging.INFO):
        self.logger = logging.getLogger(name)
        self.logger.setLevel(level)
        self.logger.propagate = False
        fmt = logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s")
        if not any(type(h) is logging.StreamHandler for h in self.logger.handlers):
            ch = logging.StreamHandler()
            ch.setFormatter(fmt)
            self.logger.addHandler(ch)
        if log_file:
            # logging.getLogger(name) is a process-wide singleton: a second Logger of the same name with another file
            # (a second Tracker with its own save_path) must get its own file handler, and the old file is released
            want = os.path.abspath(log_file)
            for h in [h for h in self.logger.handlers if isinstance(h, logging.FileHandler)]:
                if os.path.abspath(h.baseFilename) != want:
                    self.logger.removeHandler(h)
                    h.close()
            if not any(isinstance(h, logging.FileHandler) for h in self.logger.handlers):
                os.makedirs(os.path.dirname(log_file) or ".", exist_ok=True)
                fh = logging.FileHandler(log_file)
                fh.setFormatter(fmt)
                self.logger.addHandler(fh)

    def debug(self, m): self.logger.debug(m)
    def info(self, m): self.logger.info(m)
    def warning(self, m): self.logger.warning(m)
    def error(self, m): self.logger.error(m)
    def critical(self, m): self.logger.critical(m)


def load_config(path, default_path=None):
    with open(path, "r") as f:
        cfg = yaml.safe_load(f)
    parent = cfg.get("inherit_from") or default_path
    if parent:
        base = load_config(parent)
        def merge(a, b):
            for k, v in b.items():
                if isinstance(v, dict) and isinstance(a.get(k), dict):
                    merge(a[k], v)
                else:
                    a[k] = v
        merge(base, cfg)
        return base
    return cfg
