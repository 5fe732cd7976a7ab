"""constant.scm:6"""
MAX_FLOAT = 999999999999.0
