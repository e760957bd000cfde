"""Import shim: `gnnepcsaft.*` names of the reference resolved to the B200 implementation.

Putting `gnn-epc-saft_b200/` on `sys.path` (instead of the reference checkout) makes the reference's own import
lines work unchanged for the hot path:
    from gnnepcsaft.train.models import PNAPCSAFT, PNApcsaftL, PnaconvsParams, ReadoutMLPParams
    from gnnepcsaft.train.utils import create_model, calc_deg
    from gnnepcsaft.configs.default import get_config
Only the path of SURVEY.md section 8 is re-exported; everything else of the reference (thermodynamics, tuner,
featurisation) stays with the reference package."""
__b200_shim__ = True
