"""YAML pipeline configuration -- the schema of the reference's config file, unchanged
(reference microaligner/pipeline_modules/config_reader.py:75-110 field list and ranges, :149-164 read_config,
:267-304 input-type detection; examples: config_examples/*.yaml).

A config the reference accepts is accepted here and yields the same attribute tree
(``config.Input.InputImagePaths``, ``config.Output.OutputDir``, ``config.RegistrationParameters.OptFlowReg.TileSize`` ...)
and the same text when pretty-printed; a config the reference rejects is rejected with the same exception type
and message.  The implementation is table-driven: FIELDS lists every key with its type and range once, and one
validator walks the table."""
import re
from pathlib import Path
from typing import Any, Dict, Optional, Tuple

import yaml

# key -> (accepted types, minimum, maximum); maximum may name another key of the same block
REG_PARAM_FIELDS: Dict[str, Tuple[tuple, Optional[int], Any]] = {
    "NumberPyramidLevels": ((int,), 0, 8),
    "NumberIterationsPerLevel": ((int,), 1, None),
    "TileSize": ((int,), 20, None),
    "Overlap": ((int,), 10, "TileSize"),
    "NumberOfWorkers": ((int,), 0, None),
    "UseFullResImage": ((bool,), None, None),
    "UseDOG": ((bool,), None, None),
}
INPUT_FIELDS = {
    "InputImagePaths": ((dict, list), None, None),
    "ReferenceCycle": ((int,), 1, None),
    "ReferenceChannel": ((str,), None, None),
}
OUTPUT_FIELDS = {
    "OutputDir": ((str,), None, None),
    "OutputPrefix": ((str,), None, None),
    "SaveOutputToCycleStack": ((bool,), None, None),
}
_CYCLE_NAME = re.compile(r"Cycle \d+")


def _validate(block: dict, fields: Dict[str, tuple]):
    """Types first, then ranges -- the order in which the reference reports problems."""
    for name, (types, _, _) in fields.items():
        if name not in block:
            raise KeyError(f"Field {name} is absent")
        if not isinstance(block[name], types):
            expected = list(types) if len(types) == 1 else types     # the reference prints [int] but (dict, list)
            raise TypeError(f"Field {name} has wrong data type {type(block[name])}, expected {expected}")
    for name, (_, lo, hi) in fields.items():
        value = block[name]
        if isinstance(hi, str):
            hi = block[hi]
        if isinstance(value, (int, float)):
            if lo is not None and value < lo:
                raise ValueError(f"Field {name} value is smaller than minimum: {lo}")
            if hi is not None and value > hi:
                raise ValueError(f"Field {name} value is greater than maximum: {hi}")


class _Block:
    """Attribute bag that prints like the reference's config classes (``str(self.__dict__)``)."""

    def __repr__(self):
        return str(self.__dict__)


class RegParam(_Block):
    def read_from_dict(self, d: dict):
        _validate(d, REG_PARAM_FIELDS)
        for name in REG_PARAM_FIELDS:
            setattr(self, name, d[name])
        return self


class PipelineInput(_Block):
    pass


class PipelineOutput(_Block):
    pass


class PipelineRegParam:
    FeatureReg: Optional[RegParam] = None
    OptFlowReg: Optional[RegParam] = None

    def __repr__(self):
        return f"FeatureReg: {self.FeatureReg}, OptFlowReg: {self.OptFlowReg}"


class PipelineConfig(_Block):
    pass


def input_type(paths) -> str:
    """CycleStack | CycleBuilder | CyclePerImage from the shape of Input.InputImagePaths."""
    if "CycleStack" in paths:
        if len(paths) > 1:
            raise ValueError("When input is CycleStack you can specify at most 1 image path")
        return "CycleStack"
    n_dict = sum(isinstance(v, dict) for v in paths.values())
    n_str = sum(isinstance(v, str) for v in paths.values())
    if n_dict and n_str:
        raise NotImplementedError("Mixed input is not yet supported")
    if not n_dict and not n_str:
        raise ValueError("Cannot recognize type of InputImagePaths.Please check your config file against the reference.")
    if n_dict < 2 and n_str < 2:
        raise ValueError("Not enough cycles for registration. Please provide at least two cycles")
    return "CycleBuilder" if n_dict else "CyclePerImage"


def _cycle_id(name: str) -> int:
    if not _CYCLE_NAME.match(name):
        raise ValueError("Cycle names in config file should follow pattern Cycle N")
    return int(re.search(r"(\d+)", name).group(1))


def parse_paths(paths: dict, kind: str) -> dict:
    """{cycle id: Path} (CyclePerImage), {cycle id: {channel: Path}} (CycleBuilder) or {0: Path} (CycleStack)."""
    if kind == "CycleStack":
        return {0: Path(paths["CycleStack"])}
    out = {}
    for name, value in paths.items():
        cyc = _cycle_id(name)
        if kind == "CycleBuilder":
            out.setdefault(cyc, {}).update({ch: Path(p) for ch, p in value.items()})
        else:
            out[cyc] = Path(value)
    return out


class PipelineConfigReader:
    def read_config(self, config_path) -> PipelineConfig:
        with open(config_path, "r", encoding="utf-8") as f:
            raw = yaml.safe_load(f)
        missing = [k for k in ("Input", "Output", "RegistrationParameters") if k not in raw]
        if missing:
            raise ValueError("Incorrectly formatted config file.These fields are absent: " + str(missing))
        cfg = PipelineConfig()
        cfg.Input = self.parse_input(raw["Input"])
        cfg.Output = self.parse_output(raw["Output"])
        cfg.RegistrationParameters = self.parse_reg_param(raw["RegistrationParameters"])
        return cfg

    @staticmethod
    def parse_input(block) -> PipelineInput:
        if not isinstance(block, dict):
            raise ValueError("Input field is incorrect")
        _validate(block, INPUT_FIELDS)
        kind = input_type(block["InputImagePaths"])
        inp = PipelineInput()
        inp.InputImagePaths = parse_paths(block["InputImagePaths"], kind)
        inp.ReferenceCycle = block["ReferenceCycle"]
        inp.ReferenceChannel = block["ReferenceChannel"]
        inp.PipelineInputType = kind
        return inp

    @staticmethod
    def parse_output(block: dict) -> PipelineOutput:
        _validate(block, OUTPUT_FIELDS)
        out = PipelineOutput()
        out.OutputDir = Path(block["OutputDir"])
        out.OutputPrefix = block["OutputPrefix"]
        out.SaveOutputToCycleStack = block["SaveOutputToCycleStack"]
        return out

    @staticmethod
    def parse_reg_param(block: dict) -> PipelineRegParam:
        if "FeatureReg" not in block and "OptFlowReg" not in block:
            raise ValueError("Parameters for hte registration methods are absent. At least one of the registration methods: "
                             "FeatureReg or OptFlowReg must be present.")
        reg = PipelineRegParam()
        for key in ("FeatureReg", "OptFlowReg"):
            if key in block:
                _validate(block, {key: ((dict,), None, None)})
                setattr(reg, key, RegParam().read_from_dict(block[key]))
            else:
                setattr(reg, key, None)
        return reg
