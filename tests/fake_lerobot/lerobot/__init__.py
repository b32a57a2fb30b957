"""Minimal stand-in for LeRobot 0.4.x: only the symbols vla_fastvlm.lerobot_fastvla imports
(reference modeling_fastvla.py:10-12, configuration_fastvla.py:5-8, processor_fastvla.py:7-17).
TEST INFRASTRUCTURE — lets the plugin surface be imported and exercised offline."""
