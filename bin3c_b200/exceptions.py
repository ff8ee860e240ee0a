"""The exception types the hot path can raise, named as in the reference's mzd/exceptions.py."""


class ApplicationException(Exception):
    def __init__(self, message):
        super(ApplicationException, self).__init__(message)


class NoneAcceptedException(ApplicationException):
    """All sequences were excluded during filtering (raised at contact_map.py:926-927)"""
    def __init__(self):
        super(NoneAcceptedException, self).__init__('all sequences were excluded')


class UnknownEnzymeException(ApplicationException):
    """An enzyme name that is not in the table (raised by seq_utils.py:119-131 in the reference)"""
    def __init__(self, target, similar):
        super(UnknownEnzymeException, self).__init__(
            '{} is undefined, but its similar to: {}'.format(target, ', '.join(similar)))


class ParsingError(ApplicationException):
    """An error during input parsing (raised at contact_map.py:571-573)"""
    def __init__(self, msg):
        super(ParsingError, self).__init__(msg)


class ZeroLengthException(ApplicationException):
    """Sequence of zero length (raised by ExtentGrouping, contact_map.py:128-129)"""
    def __init__(self, seq_name):
        super(ZeroLengthException, self).__init__('Sequence [{}] has zero length'.format(seq_name))
