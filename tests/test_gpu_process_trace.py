"""GPU: the trace consumers (MAP table, topology report + archive) on the device store reproduce, byte for byte,
the files the reference wrote from the same trace (tests/golden/trace_mixing*)."""

import numpy as np
import pytest

from tests.test_process_trace_cpu import check_consumers

pytestmark = pytest.mark.gpu


def test_consumers_reproduce_reference_outputs_on_device(tmp_path):
    from phyclone_b200.tiles import TileStore

    check_consumers(tmp_path, lambda results: TileStore(np.array([dp.value for dp in results[0]["data"]])))
