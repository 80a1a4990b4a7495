"""Minimal object model of the host API mirror.

The reference routes every parameter through `ClassWithOptimizableVariables`
(core/base.py:30-89) and `FloatOptimizableVariable` (core/
optimizable_variable.py:232-383): values are read with `v()` / `v.evaluate()`
AT TRACE TIME, because optimisers mutate them between traces.  This module keeps
exactly that contract (annotations dict, structure dict -> attributes, `.p()`
constructors, callable variables) and nothing of the optimiser / serialiser
machinery, which sits outside the seqtrace path.
"""
import logging
import uuid


class FixedState(object):
    """Accepted for signature compatibility: FloatOptimizableVariable(FixedState(v))."""
    def __init__(self, value):
        self.value = value


class FloatOptimizableVariable(object):
    def __init__(self, state_or_value=0.0, name=""):
        value = getattr(state_or_value, "value", state_or_value)
        self._value = float(value)
        self.name = name

    def evaluate(self):
        return self._value

    __call__ = evaluate

    def setvalue(self, value):
        self._value = float(value)

    set_value = setvalue

    def __repr__(self):
        return "FloatOptimizableVariable(%r, name=%r)" % (self._value, self.name)


class ClassWithOptimizableVariables(object):
    def __init__(self, annotations_dict=None, structure_dict=None, name=""):
        self.name = name if name else str(uuid.uuid4())
        self.logger = logging.getLogger(type(self).__name__)
        self.annotations = {} if annotations_dict is None else annotations_dict
        for (key, value) in (structure_dict or {}).items():
            if not hasattr(self, key):
                setattr(self, key, value)
        self.setKind()
        self.initialize_from_annotations()

    def set_name(self, name):
        self.name = name

    def initialize_from_annotations(self):
        pass

    def setKind(self):
        self.kind = "classwithoptimizablevariables"

    # BaseLogger surface used on the path (core/log.py:39-199)
    def debug(self, msg):
        self.logger.debug(msg)

    def info(self, msg):
        self.logger.info(msg)

    def warning(self, msg):
        self.logger.warning(msg)

    def error(self, msg):
        self.logger.error(msg)
