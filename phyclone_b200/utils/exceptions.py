"""Errors raised by the data layer; class name and message text are part of the drop-in surface
(reference utils/exceptions.py)."""


class MajorCopyNumberError(Exception):
    """A genotype whose major copy number is below its minor copy number (data/pyclone.py:269-307 rejects it)."""

    _TEXT = "Major copy number should not be smaller than minor copy number\n Major_CN: %s, Minor CN: %s"

    def __init__(self, major_cn, minor_cn):
        self.major_cn, self.minor_cn = major_cn, minor_cn
        Exception.__init__(self, self._TEXT % (major_cn, minor_cn))
