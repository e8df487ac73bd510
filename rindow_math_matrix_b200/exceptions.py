"""PHP SPL exception names used by the reference drivers, as Python classes, so the parity tests
can assert the same exception type + message the PHPUnit suite asserts."""


class InvalidArgumentException(ValueError):
    pass


class RuntimeException(RuntimeError):
    pass


class LogicException(Exception):
    pass


class OutOfRangeException(IndexError):
    pass
