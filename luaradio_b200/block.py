"""Block model: type signatures, differentiate, rates.  Mirrors radio/core/block.lua:238-390,634-662.

A block is constructed with `BlockClass(args...)` (instantiate), bound to concrete input types with
`differentiate([types])`, set up with `initialize()` (may call get_rate()/get_input_type()), then
driven by `process(vec...) -> vec...` once per input vector, carrying its streaming state.
"""
from .types import Vector


class Input:
    def __init__(self, name, data_type):
        self.name, self.data_type = name, data_type


class Output:
    def __init__(self, name, data_type):
        self.name, self.data_type = name, data_type


class Port:
    aliased = False           # True for a CompositeBlock's own (aliased) ports

    def __init__(self, owner, name, kind="in"):
        self.owner, self.name, self.kind = owner, name, kind
        self.data_type = None
        self.pipe = None      # InputPort: the Pipe feeding it
        self.pipes = []       # OutputPort: the Pipes it feeds


class Pipe:
    """An edge output-port -> input-port.  In the GPU scheduler it carries no bytes itself (vectors are
    handed over in process order); it exists for rate/type propagation (radio/core/pipe.lua:27-51)."""

    def __init__(self, output, input_):
        self.output, self.input = output, input_

    def get_rate(self):
        return self.output.owner.get_rate()

    def get_data_type(self):
        return self.output.data_type


class Block:
    name = "Block"

    def __init__(self, *args, **kwargs):
        self.signatures = []
        self.inputs = None
        self.outputs = None
        self.signature = None
        self.instantiate(*args, **kwargs)

    # -- radio/core/block.lua:238-288
    def add_type_signature(self, inputs, outputs, process_func=None, initialize_func=None):
        for i, d in enumerate(inputs):
            assert isinstance(d, Input), "Invalid input port descriptor (index %d)." % (i + 1)
        for i, d in enumerate(outputs):
            assert isinstance(d, Output), "Invalid output port descriptor (index %d)." % (i + 1)
        if self.inputs is None:
            self.inputs = [Port(self, d.name, "in") for d in inputs]
        else:
            assert len(self.inputs) == len(inputs), "Invalid type signature, input count mismatch (got %d, expected %d)." % (len(inputs), len(self.inputs))
            for i, d in enumerate(inputs):
                assert self.inputs[i].name == d.name, "Invalid type signature, input name mismatch (index %d)." % (i + 1)
        if self.outputs is None:
            self.outputs = [Port(self, d.name, "out") for d in outputs]
        else:
            assert len(self.outputs) == len(outputs), "Invalid type signature, output count mismatch (got %d, expected %d)." % (len(outputs), len(self.outputs))
            for i, d in enumerate(outputs):
                assert self.outputs[i].name == d.name, "Invalid type signature, output name mismatch (index %d)." % (i + 1)
        self.signatures.append({"inputs": inputs, "outputs": outputs, "process_func": process_func, "initialize_func": initialize_func})

    # -- radio/core/block.lua:296-345
    def differentiate(self, input_data_types):
        candidates = []
        for sig in self.signatures:
            ok = True
            for i, d in enumerate(sig["inputs"]):
                if callable(d.data_type) and not hasattr(d.data_type, "type_name"):
                    pred = d.data_type(input_data_types[i])
                else:
                    pred = input_data_types[i] is d.data_type
                if not pred:
                    ok = False
                    break
            if ok:
                candidates.append(sig)
        if len(candidates) != 1:
            descs = ['"%s": [%s]' % (self.signatures[0]["inputs"][i].name, getattr(t, "type_name", "Unknown Type")) for i, t in enumerate(input_data_types)]
            raise AssertionError("No compatible type signatures found for block %s with input data types: %s." % (self.name, ", ".join(descs)))
        self.signature = candidates[0]
        if self.signature["initialize_func"] is not None:
            self.initialize = self.signature["initialize_func"].__get__(self)
        if self.signature["process_func"] is not None:
            self.process = self.signature["process_func"].__get__(self)
        for i, t in enumerate(input_data_types):
            self.inputs[i].data_type = t
        for i, d in enumerate(self.signature["outputs"]):
            self.outputs[i].data_type = input_data_types[i] if d.data_type == "copy" else d.data_type

    def get_input_type(self, index=1):
        assert self.signature, "Block not yet differentiated."
        return self.inputs[index - 1].data_type if index <= len(self.inputs) else None

    def get_output_type(self, index=1):
        assert self.signature, "Block not yet differentiated."
        return self.outputs[index - 1].data_type if index <= len(self.outputs) else None

    # -- radio/core/block.lua:383-390
    def get_rate(self):
        assert self.signature, "Block not yet differentiated."
        assert len(self.inputs) > 0, "get_rate() not implemented for source %s." % self.name
        return self.inputs[0].pipe.get_rate()

    def instantiate(self, *args):
        pass

    def initialize(self):
        pass

    def process(self, *vectors):
        raise NotImplementedError("process() not implemented")

    def cleanup(self):
        pass

    def __str__(self):
        return self.name


def factory(name, parent_class=None):
    """radio.block.factory(name, parent) (radio/core/block.lua:634-662): a new block class."""
    parent_class = parent_class or Block
    return type(name, (parent_class,), {"name": name})
