#!/bin/bash
# Regenerates the raw material of profiles/r02_sass_excerpt.txt (needs only the CUDA toolkit, no GPU).
cuobjdump -sass "$(dirname "$0")/../drone_2d_custom_gym_env_for_reinforcement_learning_b200/csrc/libdrone2d.so" | grep -E "Function :|UTCHMMA|UTCBAR|LDTM|STTM|UBLKCP|UTCATOMSWS"
