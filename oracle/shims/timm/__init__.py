"""Oracle-only stand-in for the `timm` package (absent from this image, no network).

TEST INFRASTRUCTURE: lets /root/reference import in the authoring container so that
golden vectors can be generated.  Semantics restated from SURVEY.md Appendix C.
"""
