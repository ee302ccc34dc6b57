"""Simulation calendar (host-side scalar inputs of the step kernels).

Behavioural mirror of the reference's ``grad_june/timer.py``: the clock advances by the duration
of the *current* shift, the shift index restarts when the calendar day changes, weekday/weekend
select the activity and duration tables, and the activities of a shift are ordered by the fixed
``activity_hierarchy`` (timer.py:14-26) — the order in which the kernels accumulate pressure.
"""
import calendar
import datetime as _dt

import yaml

from .paths import ensure_default_config

SECONDS_PER_DAY = 86400

activity_hierarchy = [
    "school", "university", "company", "care_home", "pub", "gym",
    "grocery", "visit", "care_visit", "cinema", "household",
]


def _as_table(x):
    """YAML gives {0: ..., 1: ...}; python callers give tuples/lists. Index by shift either way."""
    return x


class Timer:
    def __init__(
        self,
        initial_day="2020-03-01",
        total_days=10,
        weekday_step_duration=(12, 12),
        weekend_step_duration=(24,),
        weekday_activities=(("school", "household"), ("pub", "household")),
        weekend_activities=(("household",),),
    ):
        y, m, d = (int(v) for v in initial_day.split("-"))
        self.initial_date = _dt.datetime(y, m, d)
        self.total_days = total_days
        self.weekday_step_duration = _as_table(weekday_step_duration)
        self.weekend_step_duration = _as_table(weekend_step_duration)
        self.weekday_activities = _as_table(weekday_activities)
        self.weekend_activities = _as_table(weekend_activities)
        self.final_date = self.initial_date + _dt.timedelta(days=total_days)
        self.n_timesteps = 0
        self.reset()

    # -- construction ----------------------------------------------------------------
    @classmethod
    def from_file(cls, fpath=None):
        with open(fpath or ensure_default_config(), "r") as f:
            return cls.from_parameters(yaml.safe_load(f))

    @classmethod
    def from_parameters(cls, params):
        cfg = params["timer"]
        return cls(
            initial_day=cfg["initial_day"],
            total_days=cfg["total_days"],
            weekday_step_duration=cfg["step_duration"]["weekday"],
            weekend_step_duration=cfg["step_duration"]["weekend"],
            weekday_activities=cfg["step_activities"]["weekday"],
            weekend_activities=cfg["step_activities"]["weekend"],
        )

    # -- clock -----------------------------------------------------------------------
    def reset(self):
        self.date = self.initial_date
        self.previous_date = self.initial_date
        self.shift = 0
        self.delta_time = _dt.timedelta(hours=self.shift_duration)

    def __next__(self):
        self.previous_date = self.date
        self.date = self.date + self.delta_time
        self.shift = 0 if self.date.day != self.previous_date.day else self.shift + 1
        self.delta_time = _dt.timedelta(hours=self.shift_duration)
        self.n_timesteps += 1
        return self.date

    # -- derived quantities -------------------------------------------------------------
    @property
    def is_weekend(self):
        return self.date.weekday() >= 5

    @property
    def day_type(self):
        return "weekend" if self.is_weekend else "weekday"

    @property
    def now(self):
        return (self.date - self.initial_date).total_seconds() / SECONDS_PER_DAY

    @property
    def duration(self):
        return self.delta_time.total_seconds() / SECONDS_PER_DAY

    @property
    def day(self):
        return int(self.now)

    @property
    def date_str(self):
        return self.date.date().strftime("%Y-%m-%d")

    @property
    def day_of_week(self):
        return calendar.day_name[self.date.weekday()]

    @property
    def activities(self):
        return getattr(self, self.day_type + "_activities")[self.shift]

    @property
    def shift_duration(self):
        return getattr(self, self.day_type + "_step_duration")[self.shift]

    def _apply_activity_hierarchy(self, activities):
        activities.sort(key=lambda name: activity_hierarchy.index(name))
        return activities

    def get_activity_order(self):
        return self._apply_activity_hierarchy(list(self.activities))
