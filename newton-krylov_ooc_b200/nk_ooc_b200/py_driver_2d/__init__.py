"""py_driver_2d model (depth x ypos tracer transport) on the B200 path"""
