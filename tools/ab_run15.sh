#!/bin/bash
tools/evidence.sh r02k
tools/sanitize.sh r02k
