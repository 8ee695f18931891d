import logging
def get_root_logger(*a, **k):
    return logging.getLogger("mmedit")
