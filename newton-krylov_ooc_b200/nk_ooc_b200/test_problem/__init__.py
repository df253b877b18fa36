"""test_problem model (1-D column) on the B200 path"""
