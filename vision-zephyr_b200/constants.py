"""vis_zephyr/constants.py:12-14 -- the two integers the hot path depends on."""
IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200
