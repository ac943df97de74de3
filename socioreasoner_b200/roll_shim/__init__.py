"""Mirror of the reference's operator / plugin surface for the hot path (roll.distributed.strategy.*,
roll.models.model_providers, the DataProto subset those touch).  The reference's seam is this Python class hierarchy
(SURVEY.md 8b): a maintainer drops `B200InferStrategy` / `B200SegStrategy` next to the existing strategies and adds one
branch to `roll/distributed/strategy/factory.py`.  Names, argument meaning and error behaviour follow
the cited reference files so the parity tests read like tests of the reference's own strategies."""
from .protocol import DataProto, collate_fn_to_dict_list  # noqa: F401
from .strategy import B200InferStrategy, B200SegStrategy, InferenceStrategy, create_strategy  # noqa: F401
from .model_providers import default_actor_model_provider  # noqa: F401
