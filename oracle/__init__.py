"""CPU oracle for the rotated-box geometry path -- TEST INFRASTRUCTURE ONLY.

``oracle.capi``  : ctypes binding of the plain-C restatement (``geom_oracle.c``).
``oracle.ref``   : loader for the unmodified reference compiled into ``oracle/_ref``.
Nothing under ``glenet_b200/`` imports this package.
"""
