"""TEST INFRASTRUCTURE ONLY.

Minimal stand-in for the third-party ``torch_geometric`` (pinned ==2.6.1 by the reference,
environment.yml:261), which is not installed in this image and is not part of /root/reference.
It exists so that ``oracle/ref_harness.py`` can import the UNMODIFIED reference
(`/root/reference/bioemu/src`) in the build container and mint golden vectors from it.
Only the surface listed in SURVEY.md Appendix E is provided.  Never imported by the product.
"""
