"""Date / month-subset helpers of the data pipeline (reference data/utils.py:13-232), same names and semantics.

Dates are strings in ``DataConfig.datetime_format`` (``%Y-%m-%d-%H``); month groups are lists of month numbers 1..12."""
import os
import pickle
from datetime import datetime

from dateutil.relativedelta import relativedelta

from ..configs.config import DataConfig

config = DataConfig()
DATETIME_FORMAT = config.datetime_format
YEAR = frozenset(range(1, 13))


def check_valid_format(date):
    try:
        datetime.strptime(date, DATETIME_FORMAT)
        return True
    except ValueError:
        return False


def str_to_date(date):
    return datetime.strptime(date, DATETIME_FORMAT)


def date_to_str(datetime_object):
    return datetime.strftime(datetime_object, DATETIME_FORMAT)


def get_month_idx(date):
    assert check_valid_format(date), f"Date {date} is not in a valid format"
    return str_to_date(date).month


def get_month_datetime():
    return relativedelta(months=1)


def find_group_idx(month, groups):
    """1-based index of the month group that contains ``month`` (None if no group does)."""
    for idx, group in enumerate(groups):
        if month in group:
            return idx + 1
    return None


def is_full_year(months_subset):
    return months_subset is None or set(months_subset) == YEAR


def is_group_full_year(groups):
    return groups is not None and len(groups) == 1 and set(groups[0]) == YEAR


def validate_month_subset(months_subset):
    if months_subset is None:
        return True
    assert set(months_subset).issubset(YEAR), f"months_subset {months_subset} does not contain valid months"
    return True


def validate_group_months_subset(months_subset, groups):
    """The groups must partition exactly the months of ``months_subset`` (reference data/utils.py:133-163)."""
    assert months_subset is not None or groups is not None, "months_subset and groups cannot be both None"
    flat = [m for g in groups for m in g]
    assert len(flat) <= 12, f"groups {flat} has more than 12 months"
    if months_subset is None:
        assert YEAR == set(flat), f"groups missing some months {flat} from 1 to 12"
        return
    assert set(months_subset).issubset(YEAR), f"months_subset {months_subset} does not contain valid months"
    assert len(months_subset) == len(flat), f"months_subset {months_subset} has different len than {flat}"
    assert set(months_subset) == set(flat), f"months_subset {months_subset} does not contain same numbers as groups {flat}"


def month_windows(min_date, max_date):
    """Consecutive [start, end) windows that cut [min_date, max_date) at month starts, as (datetime, datetime) pairs: the walk
    both the month-subset dataset builder and the transform fitting do (reference dataset_builder.py:310-342,
    transforms.py:155-180).  The first window ends one relativedelta month after ``min_date`` (NOT snapped to day 1 -- the
    reference's quirk when ``min_date`` is not a month start); later windows end at the first of the next month."""
    end = str_to_date(max_date)
    start = str_to_date(min_date)
    nxt = start + get_month_datetime()
    while nxt < end:
        yield start, nxt
        start = nxt
        nxt = (nxt + get_month_datetime()).replace(day=1)
    yield start, end


def save_object(obj, path, filename):
    if not filename.endswith(".pkl"):
        filename = f"{filename}.pkl"
    with open(os.path.join(path, filename), "wb") as fh:
        pickle.dump(obj, fh, pickle.HIGHEST_PROTOCOL)


def log_dataset_info(dataset, dataset_name, logger):
    logger.info(f"Dataset [{dataset.__class__.__name__} - {dataset_name}] is created.")
    logger.info(f"Created {dataset.__class__.__name__} dataset of length {len(dataset)}, containing data "
                f"from {dataset.min_date} until {dataset.max_date}")
    logger.info(f"Group structure: {dataset.get_data_names()}")
    logger.info(f"Channel count: {dataset.get_channel_count()}")
    logger.info(f"Dataset size: {len(dataset)}\n")
