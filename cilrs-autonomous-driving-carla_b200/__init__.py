"""cilrs_b200 - B200-native (sm_100a) implementation of the CILRS hot path.

Frame preprocessing -> ResNet-34 + speed encoder + 4-branch command-conditioned heads (forward and backward)
-> losses -> fused Adam -> data-parallel gradient allreduce, behind the reference's own Python surface
(`CILRS.forward(image, speed, command) -> (controls, pred_speed)`, `model_state_dict` layout).
"""
__version__ = "0.1.0"
