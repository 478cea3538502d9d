class ToTensorV2: pass
