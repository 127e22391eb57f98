"""Task constants of the humanoid ping-pong variants.

The reference reads these from hydra YAML into `self.cfg["env"][...]`
(TILT:63-107, ADOF:98-116).  Hydra/OmegaConf is out of scope; the values the
hot path consumes are captured here once, with their source lines, and become
the `PpkTask` descriptor handed to the CUDA library.

Variant names follow SURVEY.md: base, a3, tilt, nes, align, a4, adof.
"""
from dataclasses import dataclass, field, replace
from typing import Tuple

# rigid-body rows used by the observation functions
# cfg/task/HumanoidPingpongTiltG1.yaml:47 (pelvis + the 9 right-arm/paddle bodies)
BODY_IDS_PINGPONG: Tuple[int, ...] = (0, 31, 32, 33, 34, 35, 36, 37, 38, 39)
# cfg/task/HumanoidPingpongTiltNESSparse27DOFG1.yaml:57
BODY_IDS_BALANCE: Tuple[int, ...] = (0, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15, 16, 17, 21, 22, 23, 24, 25, 26, 27)

VARIANT_IDS = {"base": 0, "a3": 1, "tilt": 2, "nes": 3, "align": 4, "a4": 5, "adof": 6, "align2": 7}


@dataclass(frozen=True)
class TaskConfig:
    variant: str
    num_actors: int            # A: rows of the root-state tensor per env
    num_bodies: int            # B: rows of the rigid-body tensor per env
    num_dofs: int              # D
    num_obs: int               # width of one observation row
    obs_rows: int = 1          # A4 writes one row per humanoid -> obs_buf [N,2,94]
    humanoid_actor: Tuple[int, int] = (0, 0)   # root row of humanoid 1 / 2
    ball_actor: int = 2                        # root row of the ball
    paddle_body: Tuple[int, int] = (39, 39)    # rigid-body row of paddle 1 / 2 (TILT:168, A4:172)
    pelvis_body: int = 0                       # ADOF:206
    body_ids: Tuple[int, ...] = BODY_IDS_PINGPONG
    body_ids_2: Tuple[int, ...] = ()           # A4 second humanoid (defect D13: inferred +40)
    balance_ids: Tuple[int, ...] = ()
    max_episode_length: int = 140
    alpha: float = 0.0                 # alphaVelocityReward
    power_coefficient: float = 0.0     # powerCoefficient
    penalty: float = 0.0               # penalty
    hit_table_reward: float = 0.0      # hitTableReward
    not_hit_table_penalty: float = 0.0  # nothitTablePenalty
    cross_net_reward: float = 0.0      # crossNetRewardFloat
    die_penalty: float = 0.0           # diePenaltyFloat
    hit_paddle_reward: float = 0.0     # hitPaddleReward
    miss_paddle_penalty_coefficient: float = 0.0
    is_train: bool = True
    log_every: int = 40                # stats cadence (A3:741 20, TILT:763 40, ADOF:860 32)
    flag_names: Tuple[str, ...] = ()
    flag_reset_values: Tuple[bool, ...] = ()
    counter_names: Tuple[str, ...] = ()
    reset_dof: bool = True             # NES leaves the DOF state alone (NES:871-918)
    state_names: Tuple[str, ...] = ()  # extra per-env int64 state tensors (ALIGN def #2: last_hitter)

    @property
    def variant_id(self) -> int:
        return VARIANT_IDS[self.variant]

    def with_(self, **kw) -> "TaskConfig":
        return replace(self, **kw)


_TILT_FLAGS = ("condition_calculated", "reward_calculated", "no_bounce_before_half_mask")

# cfg/task/HumanoidPingpongTiltG1.yaml:10,15,17,19; hit-table constants per defect D9 from
# cfg/task/HumanoidPingpongTiltNoEarlyStopG1.yaml:21-23
_TILT = TaskConfig(
    variant="tilt", num_actors=3, num_bodies=42, num_dofs=7, num_obs=80,
    max_episode_length=140, alpha=50.0, power_coefficient=0.0005, penalty=-200.0,
    hit_table_reward=2000.0, not_hit_table_penalty=-1000.0, log_every=40,
    flag_names=_TILT_FLAGS, flag_reset_values=(False, False, True))

CONFIGS = {
    # BASE:122-124 (5 actors: 2 robots, table, 2 balls; 83 bodies; 52 DOF), cfg/task/HumanoidPingpongG1.yaml:10
    "base": TaskConfig(variant="base", num_actors=5, num_bodies=83, num_dofs=52, num_obs=24,
                       humanoid_actor=(0, 1), ball_actor=3, paddle_body=(39, 79),
                       body_ids=(), max_episode_length=64, log_every=0),
    # A3 with the Tilt YAML constants (defect D14)
    "a3": _TILT.with_(variant="a3", log_every=20, hit_table_reward=0.0, not_hit_table_penalty=0.0,
                      flag_names=(), flag_reset_values=()),
    "tilt": _TILT,
    # cfg/task/HumanoidPingpongTiltNoEarlyStopG1.yaml:10-23
    "nes": _TILT.with_(variant="nes", max_episode_length=170, alpha=1000.0, power_coefficient=0.002,
                       penalty=-600.0, flag_names=("paddle_condition_calculated", "missed_ball_calculated"),
                       flag_reset_values=(False, False), reset_dof=False),
    "align": _TILT.with_(variant="align", flag_names=("reward_calculated",), flag_reset_values=(False,)),
    # A4:125-127: humanoid1, humanoid2, table, ball; 82 bodies; 14 DOF; ids per defect D13
    "a4": _TILT.with_(variant="a4", num_actors=4, num_bodies=82, num_dofs=14, num_obs=94, obs_rows=2,
                      humanoid_actor=(0, 1), ball_actor=3, paddle_body=(39, 79),
                      body_ids_2=tuple(i + 40 for i in BODY_IDS_PINGPONG),
                      flag_names=_TILT_FLAGS + tuple(n + "_2" for n in _TILT_FLAGS),
                      flag_reset_values=(False, False, True, False, False, True)),
    # ALIGN's second reward definition (ALIGN:1233-1351): two humanoids on the
    # 4-actor layout of A4, one shared `reward_calculated` flag and an int64 `last_hitter` state
    # ("initialised to 2", ALIGN:1253).  The reference never ran it (defect D7); constants as ALIGN.
    "align2": _TILT.with_(variant="align2", num_actors=4, num_bodies=82, num_dofs=14, num_obs=94, obs_rows=2,
                          humanoid_actor=(0, 1), ball_actor=3, paddle_body=(39, 79),
                          body_ids_2=tuple(i + 40 for i in BODY_IDS_PINGPONG),
                          flag_names=("reward_calculated",), flag_reset_values=(False,),
                          state_names=("last_hitter",)),
    # cfg/task/HumanoidPingpongTiltNESSparse27DOFG1.yaml:10-30,56-57
    "adof": TaskConfig(variant="adof", num_actors=3, num_bodies=42, num_dofs=27, num_obs=313,
                       balance_ids=BODY_IDS_BALANCE, max_episode_length=160, alpha=3000.0,
                       power_coefficient=0.002, hit_table_reward=3000.0, not_hit_table_penalty=-1000.0,
                       cross_net_reward=1000.0, die_penalty=-3000.0, hit_paddle_reward=200.0,
                       miss_paddle_penalty_coefficient=-100.0, log_every=32,
                       flag_names=("paddle_condition_calculated", "hit_table_calculated",
                                   "die_penalty_calculated", "humanoid_die_calculated"),
                       flag_reset_values=(False, False, False, False),
                       counter_names=("closer_to_paddle_count", "hit_paddle_count", "cross_net_count",
                                      "hit_table_count", "fall_down_count")),
}


def get_config(variant: str) -> TaskConfig:
    return CONFIGS[variant]
