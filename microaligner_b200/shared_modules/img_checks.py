"""Argument validation with the reference's messages (shared_modules/img_checks.py:26-47)."""


def check_img_is_2d_grey(img, img_type: str):
    if len(img.shape) > 2:
        raise ValueError(f"Expected {img_type} input to be 2D grayscale image, "
                         f"but received {img_type} image with shape {tuple(img.shape)}")


def check_img_is_provided(img, img_type: str):
    if len(img) == 0:
        raise ValueError(f"No {img_type} image provided")


def check_img_dims_match(ref, mov):
    if tuple(ref.shape) != tuple(mov.shape):
        raise ValueError("Input images have different dimensions "
                         f"reference image shape: {tuple(ref.shape)}, moving image shape: {tuple(mov.shape)}")
