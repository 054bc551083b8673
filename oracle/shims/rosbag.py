"""TEST INFRASTRUCTURE — empty stand-in: the reference imports rosbag at utils/event_camera/event.py:6
for a loader the tracking path never calls."""
