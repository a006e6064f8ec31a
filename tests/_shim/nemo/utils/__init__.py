import logging as _logging

logging = _logging.getLogger("nemo_stub")
