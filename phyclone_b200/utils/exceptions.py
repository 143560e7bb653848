class MajorCopyNumberError(Exception):
    """Same type and message as the reference's utils/exceptions.py:1-7."""

    def __init__(self, major_cn, minor_cn):
        super().__init__(
            "Major copy number should not be smaller than minor copy number"
            "\n Major_CN: {maj}, Minor CN: {minor}".format(maj=major_cn, minor=minor_cn)
        )
