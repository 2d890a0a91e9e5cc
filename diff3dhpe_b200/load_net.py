"""Model registry with the reference's lookup name (common/nets/load_net.py:5-11)."""
from .model import ConditionalDiffusionMixSTES2SGRANDLinLift


def HPE_model(MODEL_NAME):
    models = {'ConditionalDiffusionMixSTES2SGRANDLinLift': ConditionalDiffusionMixSTES2SGRANDLinLift}
    if MODEL_NAME not in models:
        raise KeyError(f"{MODEL_NAME}: only the seq2seq model is on the B200 hot path (the s2f variant is used by no "
                       "shipped config, SURVEY.md section 2 row 5)")
    return models[MODEL_NAME]
