"""TEST INFRASTRUCTURE - write-only stand-in for the absent third-party package h5py, just enough for the
reference's ``write_calibration_file`` (bilby/gw/detector/calibration.py:152-205) to run while golden vectors are
generated: groups and datasets are accepted and discarded, nothing is written, nothing can be read back."""


class _Group:
    def __init__(self):
        self.attrs = {}

    def create_group(self, name):
        return _Group()

    def create_dataset(self, name, data=None, **kwargs):
        return None


class File(_Group):
    def __init__(self, filename, mode="r"):
        super().__init__()
        if "w" not in mode:
            raise OSError("the h5py stand-in cannot read files")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False
