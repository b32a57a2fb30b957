def policy_action_to_transition(action):
    return action


def transition_to_policy_action(transition):
    return transition
