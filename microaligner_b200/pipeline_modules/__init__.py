"""Pipeline support of the YAML entry point (reference microaligner/pipeline_modules/): configuration schema,
dataset structure (cycle -> channel -> z-plane -> TIFF page) and the OME-XML written into the output files.
No pixels are touched here; the per-cycle dispatch lives in microaligner_b200/__main__.py."""
from .config_reader import PipelineConfig, PipelineConfigReader  # noqa: F401
from .metadata_handling import DatasetStruct, DatasetStructCreator  # noqa: F401
from .ome_meta_processing import create_new_meta  # noqa: F401
