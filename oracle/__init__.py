"""ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/me_oracle.py header).

Nothing under this directory is imported by `room_instance_segmentation_mask3d_b200`.
"""
