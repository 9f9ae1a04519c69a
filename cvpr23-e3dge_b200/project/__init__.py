"""Reference-facing import paths (`project.utils.volume_renderer`,
`project.models.stylesdf_model`, `project.models.op`) re-exporting e3dge_b200, so code written
against NIRVANALAN/CVPR23-E3DGE imports the B200 implementation unchanged when
`cvpr23-e3dge_b200/` is on sys.path ahead of the reference checkout."""
