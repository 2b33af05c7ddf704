"""Empty: the reference dead code imports gym.scoreboard."""
