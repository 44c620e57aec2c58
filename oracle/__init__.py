"""Test infrastructure only: CPU restatement of the reference path (amodal_oracle) and seeded synthetic weights/inputs
(synth). Nothing under amodal-depth-anything_b200/ may import this package."""
