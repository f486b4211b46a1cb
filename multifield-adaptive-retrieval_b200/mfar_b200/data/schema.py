"""Dataset field presets and ``resolve_fields``.

Same contract as mfar/data/schema.py:96-134: a comma-separated scorer spec
("all_dense,all_sparse", "single_dense", "title_dense", ...) resolves to an ORDERED dict of
fields - dense keys sorted, then sparse keys sorted (schema.py:130-134).  That order fixes the
column order of the mixture matrix W, of the field mask and of the per-field score tensors.
The per-field token budgets are the reference's presets (schema.py:11-69).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Union

from .typedef import Field, FieldType

SPARSE_MAX = 1048576

# field name -> max token length of its dense encoder input
_PRESETS: Dict[str, Dict[str, int]] = {
    "mag": {
        "abstract": 512, "author___affiliated_with___institution": 512, "paper___cites___paper": 512,
        "paper___has_topic___field_of_study": 64, "title": 64,
    },
    "prime": {
        "associated with": 256, "carrier": 8, "contraindication": 128, "details": 512, "enzyme": 64,
        "expression absent": 64, "expression present": 512, "indication": 32, "interacts with": 512,
        "linked to": 8, "name": 64, "off-label use": 8, "parent-child": 256, "phenotype absent": 8,
        "phenotype present": 512, "ppi": 512, "side effect": 128, "source": 8,
        "synergistic interaction": 512, "target": 64, "transporter": 8, "type": 8,
    },
    "amazon": {
        "also_buy": 512, "also_view": 512, "brand": 16, "description": 512, "feature": 512, "qa": 512,
        "review": 512, "title": 128,
    },
    "whatsthatbook": {
        "author": 16, "author_url": 64, "date": 64, "description": 512, "genres": 64, "id": 16,
        "image_link": 64, "isbn_13": 16, "parsed_dates": 16, "ratings": 16, "reviews": 16, "title": 64,
    },
}
DATASET_NAMES: List[str] = list(_PRESETS)


def _schema(dataset: str) -> Dict[str, Field]:
    out: Dict[str, Field] = {}
    for name, max_len in _PRESETS[dataset].items():
        out[f"{name}_sparse"] = Field(f"{name}_sparse", name, FieldType.SPARSE, SPARSE_MAX, dataset)
        out[f"{name}_dense"] = Field(f"{name}_dense", name, FieldType.DENSE, max_len, dataset)
    return out


SCHEMAS: Dict[str, Dict[str, Field]] = {d: _schema(d) for d in DATASET_NAMES}


def _single(dataset: str, kind: FieldType) -> Field:
    if kind is FieldType.SPARSE:
        return Field("single_sparse", "single", FieldType.SPARSE, SPARSE_MAX, dataset)
    return Field("single_dense", "single", FieldType.DENSE, 512, dataset)


def _dataset_of(path_or_name: str) -> str:
    leaf = path_or_name.split("/")[-1]
    for d in DATASET_NAMES:
        if d in leaf:
            return d
    raise NotImplementedError(f"Dataset {path_or_name} is not supported!")


def resolve_fields(field_names: Union[str, Iterable[str]], dataset: str) -> Dict[str, Field]:
    name = _dataset_of(dataset)
    valid = SCHEMAS[name]
    if isinstance(field_names, str):
        field_names = [part.replace(".", " ") for part in field_names.split(",")]   # schema.py:107-109
    picked: Dict[str, Field] = {}
    for spec in field_names:
        if spec in ("all_sparse", "all_dense"):
            want = FieldType.SPARSE if spec == "all_sparse" else FieldType.DENSE
            picked.update({k: f for k, f in valid.items() if f.field_type is want})
        elif spec == "single_sparse":
            picked[spec] = _single(name, FieldType.SPARSE)
        elif spec == "single_dense":
            picked[spec] = _single(name, FieldType.DENSE)
        elif spec in valid:
            picked[spec] = valid[spec]
        else:
            raise ValueError(f"Field {spec} not found in dataset {dataset}")
    keys = sorted(picked)
    dense = [k for k in keys if picked[k].field_type is FieldType.DENSE]
    sparse = [k for k in keys if picked[k].field_type is FieldType.SPARSE]
    return {k: picked[k] for k in dense + sparse}
