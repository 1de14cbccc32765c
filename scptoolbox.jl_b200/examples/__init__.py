"""Example problem definitions (mirrors of test/examples/* of the reference) for the B200 path."""
